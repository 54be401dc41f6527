"""Small, ragged exercise of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rocket_landing_rl_b200 import RocketLandingSB3VecEnv, RocketLandingVecEnv   # noqa: E402

for n in (1, 33, 257, 5000):
    for substeps in (1, 3):
        env = RocketLandingVecEnv(n, seed=1, substeps=substeps)
        env.reset()
        g = torch.Generator(device="cuda").manual_seed(0)
        for t in range(140 // substeps):
            a = torch.rand((n, 2), device="cuda", generator=g)
            a[:, 1] = a[:, 1] * 2 - 1
            env.step(a)
        st = env.get_state()
        env.set_state(st, mask=torch.arange(n, device="cuda") % 2 == 0)
        env.reset(mask=torch.arange(n, device="cuda") % 3 == 0)
        s = env.stats()
        assert s.env_steps > 0 and s.nonfinite == 0
        env.close()
os.environ["RL_B200_GENERIC_KERNEL"] = "1"
env = RocketLandingVecEnv(777, seed=2, auto_reset=False)
env.reset()
for t in range(20):
    env.step(torch.zeros((777, 2), device="cuda"))
env.close()
del os.environ["RL_B200_GENERIC_KERNEL"]
host = RocketLandingSB3VecEnv(1000, seed=3)
host.reset()
for t in range(130):
    host.step(np.random.default_rng(t).uniform([0, -1], [1, 1], (1000, 2)).astype(np.float32))
host.close()
torch.cuda.synchronize()
print("sanitizer probe finished")
