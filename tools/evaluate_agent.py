#!/usr/bin/env python
"""Evaluate a stable-baselines3 PPO checkpoint of the reference (``assets/model/v3/best_model.zip`` +
``vecnormalize.pkl``, or the committed fixture ``tests/golden/ref_policy_v3.npz``) over many first
episodes at once: B200 environment step + device VecNormalize + fused tcgen05 policy kernel,
deterministic actions as ``backend/rl/agent.py`` takes them.  Prints one JSON line.

    python tools/evaluate_agent.py --rockets 1048576
    python tools/evaluate_agent.py --zip /path/best_model.zip --pkl /path/vecnormalize.pkl
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rocket_landing_rl_b200 import (ActorCriticMLP, DeviceVecNormalize, FusedActorCritic,   # noqa: E402
                                    RocketLandingVecEnv, read_sb3_vecnormalize)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--rockets", type=int, default=1 << 20)
    ap.add_argument("--zip", default=None)
    ap.add_argument("--pkl", default=None)
    ap.add_argument("--fixture", default=os.path.join(ROOT, "tests", "golden", "ref_policy_v3.npz"))
    args = ap.parse_args()
    if args.zip:
        net = ActorCriticMLP.from_sb3_zip(args.zip)
        st = read_sb3_vecnormalize(args.pkl)
    else:
        fx = dict(np.load(args.fixture))
        net = ActorCriticMLP.from_sb3_state_dict({k[2:].replace("__", "."): v for k, v in fx.items() if k.startswith("w_")})
        st = {k: fx[k] for k in ("obs_mean", "obs_var", "obs_count", "ret_mean", "ret_var", "ret_count", "clip_obs", "epsilon")}
    n = args.rockets
    raw = RocketLandingVecEnv(n, seed=1, auto_reset=False)
    env = DeviceVecNormalize(raw, training=False, norm_reward=False, clip_obs=float(st["clip_obs"]), epsilon=float(st["epsilon"]))
    env.load_stats({k: st[k] for k in ("obs_mean", "obs_var", "obs_count", "ret_mean", "ret_var", "ret_count")})
    policy = FusedActorCritic(net.cuda())
    low, high = torch.tensor([0.0, -1.0], device="cuda"), torch.tensor([1.0, 1.0], device="cuda")
    obs = env.reset()
    first_done = torch.zeros(n, dtype=torch.int32, device="cuda")
    outcome = torch.zeros(n, dtype=torch.int8, device="cuda")
    ret = torch.zeros(n, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(raw.max_episode_steps):
        action = torch.clamp(policy.action_mean(obs), low, high)
        obs, _, term, trunc, info = env.step(action)
        live = first_done == 0
        ret += torch.where(live, env.get_original_reward(), torch.zeros_like(ret))
        new = (term | trunc) & live
        first_done = torch.where(new, torch.full_like(first_done, t + 1), first_done)
        outcome = torch.where(new & term, info["landing"], outcome)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    freq = (torch.bincount(outcome.long(), minlength=5).float() / n).cpu().tolist()
    print(json.dumps({"episodes": n, "seconds": round(dt, 3), "env_steps_per_s": n * raw.max_episode_steps / dt,
                      "frequencies": dict(zip(("time_limit_or_out_of_bounds", "safe", "good", "ok", "unsafe"),
                                              [round(f, 4) for f in freq])),
                      "mean_episode_length": float(first_done.float().mean()), "mean_return": float(ret.mean())}))


if __name__ == "__main__":
    main()
