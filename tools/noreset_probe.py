"""Step-kernel time with and without finished episodes: with auto-reset off and every rocket
landed, no episode ever ends again and only the always-executed path runs -- the upper bound of what
moving the reset path out of the step kernel could give (profiles/README.md, "what is left").
GPU box only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocket_landing_rl_b200 import RocketLandingVecEnv
for n in (16777216, 67108864):
    for auto in (True, False):
        env = RocketLandingVecEnv(n, seed=0, auto_reset=auto)
        env.reset()
        g = torch.Generator(device="cuda").manual_seed(1)
        acts = []
        for _ in range(4):
            a = torch.rand((n, 2), device="cuda", generator=g); a[:, 1].mul_(2).sub_(1); acts.append(a)
        for t in range(600 if auto else 300):      # auto: mixed phases; no auto: every rocket has landed, no episode ever ends again
            env.step(acts[t % 4])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for t in range(100):
            env.step(acts[t % 4])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 100
        print(n, "auto_reset" if auto else "no resets (upper bound)", round(ms, 4), "ms", round(123 * n / ms / 1e6 / 6543.7, 3))
        env.close(); del env, acts; torch.cuda.empty_cache()
