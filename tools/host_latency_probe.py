#!/usr/bin/env python
"""Per-call latency of the HOST-buffer VecEnv (RocketLandingSB3VecEnv: numpy in, numpy out -- the
call scripts/train.py's PPO makes) at the batch sizes the reference itself uses (rl.n_envs = 8,
config.yaml:95) and a few larger ones."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocket_landing_rl_b200 import RocketLandingSB3VecEnv   # noqa: E402

out = {}
for n in (8, 64, 1024, 4096, 65536):
    row = {}
    for name, infos in (("step_raw", False), ("step_with_infos", True)):
        env = RocketLandingSB3VecEnv(n, seed=0, build_infos=infos)
        env.reset()
        a = np.random.default_rng(0).uniform([0, -1], [1, 1], size=(n, 2)).astype(np.float32)
        fn = (lambda: env.step_raw(a)) if not infos else (lambda: env.step(a))
        for _ in range(200):
            fn()
        reps = 2000 if n <= 4096 else 300
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        us = (time.perf_counter() - t0) / reps * 1e6
        row[name + "_us"] = round(us, 2)
        row[name + "_env_steps_per_s"] = round(n / (us * 1e-6))
        env.close()
    out[n] = row
print(json.dumps(out))
