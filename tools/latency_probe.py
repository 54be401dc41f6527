#!/usr/bin/env python
"""Per-call cost of RocketLandingVecEnv.step for small batches (launch-bound regime,
BASELINE.json configs[1]: 4096 rockets), eager and as a CUDA-graph replay."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocket_landing_rl_b200 import RocketLandingVecEnv   # noqa: E402

out = {}
for n in (8, 4096, 65536, 1 << 20):
    env = RocketLandingVecEnv(n, seed=0)
    env.reset()
    a = torch.rand((n, 2), device="cuda")
    for _ in range(200):
        env.step(a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000):
        env.step(a)
    host_us = (time.perf_counter() - t0) / 2000 * 1e6           # host-side enqueue cost
    torch.cuda.synchronize()
    wall_us = (time.perf_counter() - t0) / 2000 * 1e6
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(16):
                env.step(a)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(125):
        g.replay()
    torch.cuda.synchronize()
    graph_us = (time.perf_counter() - t0) / 2000 * 1e6
    out[n] = {"eager_enqueue_us": round(host_us, 2), "eager_us_per_step": round(wall_us, 2),
              "graph16_us_per_step": round(graph_us, 2), "graph_env_steps_per_s": n / (graph_us * 1e-6)}
    env.close()
print(json.dumps(out))
