"""Generate the committed golden vectors from the UNMODIFIED reference.

Run in the build container only (needs the reference checkout):

    python tests/golden/make_golden.py

Writes three ``.npz`` files beside this script.  Every number in them is an output of
the reference's own ``RocketLandingEnv`` (backend/envs/lander.py) imported through
``oracle/ref_loader.py`` -- not of the oracle and not of this repo's kernels.

* ``ref_rollout_single.npz``  BASELINE.json configs[0]: one env, ``np.random.seed(0)``,
  actions ``default_rng(0).uniform([0,-1],[1,1]).astype(float32)``, 1000 steps,
  ``reset()`` on done (SB3 auto-reset contract).
* ``ref_rollout_batch8.npz``  eight more such envs (seeds 100..107), stacked [T, E, ...].
* ``ref_edge_cases.npz``      teacher-forced wild states (``oracle/pin_against_reference.
  edge_case_states``), two steps each: landing classes, +-180 wraps, out of bounds,
  tip-over, empty tank, max-step truncation, dead-band reward branches.

Per step the files hold the RAW post-step state (``info["raw_state"]`` words in
``STATE_WORDS`` order: x y vx vy ax ay angle angularVelocity mass fuelMass), the three
previous-state words the integrator reads, the clipped float32 observation, reward,
flags, landing code (protocol.py:31-41 numbering) and, on done steps, the freshly
sampled reset state that the replay must inject.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, _ROOT)

from oracle.ref_loader import load_reference                      # noqa: E402
from oracle.pin_against_reference import (REF_KEYS, _inject_reference,   # noqa: E402
                                          edge_case_states)

LANDING_CODE = {"safe": 1, "good": 2, "ok": 3, "unsafe": 4}


def _vec(env):
    s = env.rocket.state
    return [float(s[k]) for k in REF_KEYS]


def _prev(env):
    p = env.rocket.previous_state
    return [float(p["x"]), float(p["y"]), float(p["angle"])]


def rollout(ref, seed: int, action_seed: int, steps: int) -> dict:
    env = ref.RocketLandingEnv()
    np.random.seed(seed)
    obs0, _ = env.reset()
    rng = np.random.default_rng(action_seed)
    out = {
        "init_state": np.array(_vec(env)), "init_obs": obs0.copy(),
        "actions": np.zeros((steps, 2), np.float32),
        "state": np.zeros((steps, 10)), "prev": np.zeros((steps, 3)),
        "obs": np.zeros((steps, 8), np.float32), "terminal_obs": np.zeros((steps, 8), np.float32),
        "reward": np.zeros(steps), "terminated": np.zeros(steps, bool),
        "truncated": np.zeros(steps, bool), "landing": np.zeros(steps, np.int8),
        "steps": np.zeros(steps, np.int32), "reset_state": np.zeros((steps, 10)),
    }
    for t in range(steps):
        a = rng.uniform([0, -1], [1, 1]).astype(np.float32)
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


def edge_cases(ref, n: int) -> dict:
    env = ref.RocketLandingEnv()
    cases = edge_case_states(np.random.default_rng(7), n)
    out = {
        "init_state": np.zeros((n, 10)), "init_step": np.zeros(n, np.int32),
        "actions": np.zeros((n, 2), np.float32),
        "state": np.zeros((2, n, 10)), "prev": np.zeros((2, n, 3)),
        "obs": np.zeros((2, n, 8), np.float32), "reward": np.zeros((2, n)),
        "terminated": np.zeros((2, n), bool), "truncated": np.zeros((2, n), bool),
        "landing": np.zeros((2, n), np.int8),
    }
    for i, (st, act, step) in enumerate(cases):
        _inject_reference(env, st, step)
        out["init_state"][i] = [float(st[k]) for k in REF_KEYS]
        out["init_step"][i], out["actions"][i] = step, act
        for k in range(2):
            obs, r, term, trunc, _ = env.step(act)
            out["state"][k, i], out["prev"][k, i], out["obs"][k, i] = _vec(env), _prev(env), obs
            out["reward"][k, i], out["terminated"][k, i], out["truncated"][k, i] = r, term, trunc
            if term:
                msg = ref.evaluate_landing(env.rocket.get_state(), env.config)["landing_message"]
                out["landing"][k, i] = LANDING_CODE[msg]
    return out


def main():
    ref = load_reference()
    single = rollout(ref, seed=0, action_seed=0, steps=1000)
    np.savez_compressed(os.path.join(_HERE, "ref_rollout_single.npz"), **single)
    runs = [rollout(ref, seed=100 + e, action_seed=100 + e, steps=1000) for e in range(8)]
    batch = {k: np.stack([r[k] for r in runs], axis=1 if runs[0][k].ndim >= 1 and
                         k not in ("init_state", "init_obs") else 0) for k in runs[0]}
    np.savez_compressed(os.path.join(_HERE, "ref_rollout_batch8.npz"), **batch)
    edge = edge_cases(ref, 3000)
    np.savez_compressed(os.path.join(_HERE, "ref_edge_cases.npz"), **edge)
    for f in ("ref_rollout_single.npz", "ref_rollout_batch8.npz", "ref_edge_cases.npz"):
        print(f, os.path.getsize(os.path.join(_HERE, f)), "bytes")
    print("single: episodes", int((single["terminated"] | single["truncated"]).sum()),
          "| edge: terminated", int(edge["terminated"].sum()), "truncated",
          int(edge["truncated"].sum()), "landing classes",
          np.bincount(edge["landing"].ravel(), minlength=5).tolist())


if __name__ == "__main__":
    main()
