import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from rocket_landing_rl_b200 import RocketLandingEnv
env = RocketLandingEnv(seed=1)
env.reset()
a = np.array([0.6, 0.1], np.float32)
for _ in range(50): env.step(a)
env.reset()
t0 = time.perf_counter(); n = 0
for _ in range(2000):
    o, r, te, tr, info = env.step(a); n += 1
    if te or tr: env.reset()
print("RocketLandingEnv.step: %.1f us per call" % ((time.perf_counter() - t0) / n * 1e6))
