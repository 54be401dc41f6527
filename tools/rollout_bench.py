#!/usr/bin/env python
"""BASELINE.json configs[4]: end-to-end PPO rollout collection with GPU-resident rockets feeding
on-device buffers (envs -> VecNormalize -> policy MLP -> buffer -> GAE), one GPU.  Prints one JSON
line with the env-steps/s of the whole collection loop and of its parts.  Not the bench headline
(bench.py is); kept as the measurement of the 'next' rows of SURVEY.md section 8f."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rocket_landing_rl_b200 import (ActorCriticMLP, DeviceRolloutBuffer, DeviceVecNormalize,   # noqa: E402
                                    FusedActorCritic, RocketLandingVecEnv, collect_rollout)


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rockets", type=int, default=1 << 20)
    ap.add_argument("--n-steps", type=int, default=32)
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--dtype", default="tf32", choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--fused", action="store_true", help="evaluate the policy with rl_mlp_forward (csrc/rl_policy.cu)")
    ap.add_argument("--no-bootstrap", action="store_true",
                    help="diagnosis: skip the time-limit bootstrap (no flag kernel, no event per step)")
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = args.dtype != "fp32"
    torch.backends.cudnn.allow_tf32 = args.dtype != "fp32"
    n, T = args.rockets, args.n_steps
    raw = RocketLandingVecEnv(n, seed=0)
    env = DeviceVecNormalize(raw, gamma=0.99)
    policy = ActorCriticMLP().cuda()
    if args.fused:
        policy = FusedActorCritic(policy)
    buf = DeviceRolloutBuffer(T, n, "cuda")
    obs = env.reset()
    starts = torch.ones(n, dtype=torch.bool, device="cuda")
    state = {"obs": obs, "starts": starts}

    def one_pass():
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if args.dtype == "bf16" else torch.autocast("cuda", enabled=False)
        with ctx:
            state["obs"], state["starts"] = collect_rollout(env, policy, buf, state["obs"], state["starts"],
                                                            bootstrap_timeouts=not args.no_bootstrap)

    for _ in range(4):           # ~128 steps: first episodes end, buffers warm
        one_pass()
    ms_pass = timed(one_pass, args.passes)
    a = torch.rand((n, 2), device="cuda")
    ms_env = timed(lambda: raw.step(a), 50)
    ms_envnorm = timed(lambda: env.step(a), 50)
    with torch.no_grad():
        for _ in range(3):       # (the output tensors of a stand-alone forward are new sizes for the caching allocator)
            policy(state["obs"])
        ms_policy = timed(lambda: policy(state["obs"]), 20)
    ms_gae = timed(lambda: buf.compute_returns_and_advantage(torch.zeros(n, device="cuda"), state["starts"], 0.99, 0.95), 20)
    print(json.dumps({
        "workload": f"{n} rockets, n_steps {T}, policy 8-256-256 x2 (136965 params, "
                    + ("fused tensor-core kernel, bf16 weights)" if args.fused else f"{args.dtype} GEMMs)"),
        "rollout_env_steps_per_s": n * T / (ms_pass * 1e-3), "ms_per_collection_pass": ms_pass,
        "ms_env_step": ms_env, "ms_env_step_plus_vecnormalize": ms_envnorm, "ms_policy_forward": ms_policy,
        "ms_gae_T_steps": ms_gae, "env_only_steps_per_s": n / (ms_env * 1e-3),
        "env_plus_vecnormalize_steps_per_s": n / (ms_envnorm * 1e-3),
        "gae_GBps": 17.0 * n * T / (ms_gae * 1e-3) / 1e9,
    }))


if __name__ == "__main__":
    main()
