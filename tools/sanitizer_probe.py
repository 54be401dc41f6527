#!/usr/bin/env python
"""Small workload for `compute-sanitizer --tool memcheck|racecheck|initcheck python tools/sanitizer_probe.py`:
4096 rockets (ragged: 4001), reset, 130 steps with auto-reset (every rocket finishes an episode), fused
substeps, the optional outputs, set/get state, statistics and the host path."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocket_landing_rl_b200 import RocketLandingSB3VecEnv, RocketLandingVecEnv   # noqa: E402

n = 4001
env = RocketLandingVecEnv(n, seed=1, raw_state=True)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
for t in range(130):
    a = torch.rand((n, 2), device="cuda", generator=g)
    a[:, 1] = a[:, 1] * 2 - 1
    env.step(a)
env.substeps = 4
for t in range(10):
    env.step(a)
env.set_state(env.get_state())
st = env.stats()
assert st.episodes > 0.9 * n and st.nonfinite == 0, st
host = RocketLandingSB3VecEnv(n, seed=2)
host.reset()
for t in range(20):
    host.step(np.random.default_rng(t).uniform([0, -1], [1, 1], (n, 2)).astype(np.float32))
torch.cuda.synchronize()
print("sanitizer probe OK:", int(st.env_steps), "env steps,", int(st.episodes), "episodes")
