"""Long-horizon golden trajectories from the UNMODIFIED reference: episodes that run to the
1000-step limit (``lander.py:237-241``), which random actions never reach (they crash after
91-154 steps).

Run in the build container only (needs the reference checkout):

    python tests/golden/make_long_horizon.py

Two drivers, both flying the live ``RocketLandingEnv`` (through ``oracle/ref_loader.py``):

* ``ref_long_agent_v3.npz``  the reference's own shipped agent (``assets/model/v3``), driven as
  ``backend/rl/agent.py:160-205`` drives it (normalise with the pickled VecNormalize statistics,
  deterministic action = policy mean, clip to the Box).  60 % of its episodes hit the time limit.
* ``ref_long_hover.npz``     a hover controller (holds ~300 m with a PD attitude loop and 2 % action
  noise) that keeps EVERY rocket aloft to the limit -- the worst case for accumulated error.

Same per-step record as ``make_golden.py`` ([T, E, ...]: actions, raw post-step state, previous
x / y / angle, observation, terminal observation, reward, flags, landing code, reset state), so
``tests/parity.replay_golden_rollout`` replays them word by word.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
sys.path[:0] = [_ROOT, _HERE]

from oracle.ref_loader import load_reference                      # noqa: E402
from make_golden import LANDING_CODE, _prev, _vec                 # noqa: E402
from make_policy_fixture import load_checkpoint                   # noqa: E402


def agent_policy(version: str = "v3"):
    sd, st = load_checkpoint(version)
    w = {k: v.numpy().astype(np.float64) for k, v in sd.items()}
    low, high = np.array([0.0, -1.0]), np.array([1.0, 1.0])

    def act(env, obs, rng):
        norm = np.clip((obs.astype(np.float64) - st["obs_mean"]) / np.sqrt(st["obs_var"] + st["epsilon"]),
                       -st["clip_obs"], st["clip_obs"]).astype(np.float32).astype(np.float64)
        h = np.tanh(norm @ w["mlp_extractor.policy_net.0.weight"].T + w["mlp_extractor.policy_net.0.bias"])
        h = np.tanh(h @ w["mlp_extractor.policy_net.2.weight"].T + w["mlp_extractor.policy_net.2.bias"])
        mean = h @ w["action_net.weight"].T + w["action_net.bias"]
        return np.clip(mean, low, high).astype(np.float32)
    return act


def hover_policy(env, obs, rng):
    s = env.rocket.state
    y, vy, a, w = s["y"], s["vy"], s["angle"], s["angularVelocity"]
    m = s["mass"] + s["fuelMass"]
    tv = -40.0 if y > 400 else float(np.clip(-(y - 300) * 0.2, -20, 20))
    th = (9.81 * m / 1e7) * (1 + float(np.clip((tv - vy) * 0.3, -1, 6)))
    cg = float(np.clip(-(a * 0.9 + w * 1.6), -1, 1))
    return (np.array([th, cg]) + rng.normal(0, 0.02, 2)).astype(np.float32)


def rollout(ref, policy, seed: int, steps: int) -> dict:
    env = ref.RocketLandingEnv()
    np.random.seed(seed)
    obs, _ = env.reset()
    rng = np.random.default_rng(seed)
    out = {
        "init_state": np.array(_vec(env)), "init_obs": obs.copy(),
        "actions": np.zeros((steps, 2), np.float32),
        "state": np.zeros((steps, 10)), "prev": np.zeros((steps, 3)),
        "obs": np.zeros((steps, 8), np.float32), "terminal_obs": np.zeros((steps, 8), np.float32),
        "reward": np.zeros(steps), "terminated": np.zeros(steps, bool),
        "truncated": np.zeros(steps, bool), "landing": np.zeros(steps, np.int8),
        "steps": np.zeros(steps, np.int32), "reset_state": np.zeros((steps, 10)),
    }
    for t in range(steps):
        a = policy(env, obs, rng)
        obs, r, term, trunc, info = env.step(a)
        out["actions"][t] = a
        out["state"][t], out["prev"][t] = _vec(env), _prev(env)
        out["reward"][t], out["terminated"][t], out["truncated"][t] = r, term, trunc
        out["steps"][t] = info["steps"]
        if term:
            msg = ref.evaluate_landing(env.rocket.get_state(), env.config)["landing_message"]
            out["landing"][t] = LANDING_CODE[msg]
        if term or trunc:                       # SB3 VecEnv auto-reset contract
            out["terminal_obs"][t] = obs
            obs, _ = env.reset()
            out["reset_state"][t] = _vec(env)
        out["obs"][t] = obs
    return out


def make(ref, name: str, policy, envs: int, steps: int, seed0: int) -> None:
    runs = [rollout(ref, policy, seed0 + e, steps) for e in range(envs)]
    batch = {k: np.stack([r[k] for r in runs], axis=1 if k not in ("init_state", "init_obs") else 0)
             for k in runs[0]}
    path = os.path.join(_HERE, name)
    np.savez_compressed(path, **batch)
    done = batch["terminated"] | batch["truncated"]
    lengths = batch["steps"][done]
    print(f"{name}: {os.path.getsize(path)} bytes, {int(done.sum())} episodes, lengths {sorted(lengths.tolist())}, "
          f"time-limit {int((batch['truncated'] & ~batch['terminated']).sum())}, landings "
          f"{np.bincount(batch['landing'].ravel(), minlength=5)[1:].tolist()}")


def main():
    ref = load_reference()
    make(ref, "ref_long_agent_v3.npz", agent_policy("v3"), envs=8, steps=1100, seed0=300)
    make(ref, "ref_long_hover.npz", hover_policy, envs=8, steps=1100, seed0=400)


if __name__ == "__main__":
    main()
