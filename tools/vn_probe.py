"""Six env + VecNormalize steps at 16 Mi rockets, for `ncu -k regex:vn_` launch lists
(profiles/README.md, VecNormalize kernels).  GPU box only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocket_landing_rl_b200 import DeviceVecNormalize, RocketLandingVecEnv
n = 16777216
raw = RocketLandingVecEnv(n, seed=0)
env = DeviceVecNormalize(raw, gamma=0.99)
env.reset()
a = torch.rand((n, 2), device="cuda")
for _ in range(6):
    env.step(a)
torch.cuda.synchronize()
