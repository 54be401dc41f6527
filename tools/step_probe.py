#!/usr/bin/env python
"""Quick A/B probe of the step kernel: env-steps/s of a steady-state run of one library build.

    RL_B200_LIBRARY=variants/lib_x.so python tools/step_probe.py [--rockets N] [--substeps K] [--launches L]

Spins up 2000 env steps (episodes desynchronised), then times L back-to-back launches with CUDA
events, `--repeat` times.  Under `ncu --metrics smsp__inst_executed.sum,... -k regex:step_kernel
--launch-skip S --launch-count 1` it gives the dynamic instruction count of one steady-state launch."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocket_landing_rl_b200 import RocketLandingVecEnv   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rockets", type=int, default=64 * 1024 * 1024)
    ap.add_argument("--substeps", type=int, default=1)
    ap.add_argument("--launches", type=int, default=60)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--spinup", type=int, default=2000)
    ap.add_argument("--tag", default=os.environ.get("RL_B200_LIBRARY", "in-tree"))
    a = ap.parse_args()
    n, K = a.rockets, a.substeps
    env = RocketLandingVecEnv(n, seed=0, substeps=K)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(1234)
    acts = []
    for _ in range(4):
        t = torch.rand((n, 2), device="cuda", generator=gen)
        t[:, 1].mul_(2.0).sub_(1.0)
        acts.append(t)
    for t in range((a.spinup + K - 1) // K):
        env.step(acts[t % 4])
    env.stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rates = []
    for _ in range(a.repeat):
        torch.cuda.synchronize()
        e0.record()
        for t in range(a.launches):
            env.step(acts[t % 4])
        e1.record()
        torch.cuda.synchronize()
        st = env.stats(reset=True)
        rates.append(st.env_steps / (e0.elapsed_time(e1) * 1e-3))
    print(json.dumps({"tag": os.path.basename(a.tag), "rockets": n, "substeps": K,
                      "env_steps_per_s": [float("%.4g" % r) for r in rates],
                      "frac_123B": [round(123.0 / K * r / 6543.7e9, 4) for r in rates]}), flush=True)
    env.close()


if __name__ == "__main__":
    main()
