import sys, torch
sys.path.insert(0, '/root/repo')
from rocket_landing_rl_b200 import DeviceVecNormalize, RocketLandingVecEnv
n = 16777216
raw = RocketLandingVecEnv(n, seed=0)
env = DeviceVecNormalize(raw, gamma=0.99)
env.reset()
a = torch.rand((n, 2), device="cuda")
for _ in range(6):
    env.step(a)
torch.cuda.synchronize()
