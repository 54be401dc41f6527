"""Pin ``oracle/rocket_oracle.py`` against the UNMODIFIED reference, bit for bit.

TEST INFRASTRUCTURE ONLY.  Runs only where the reference checkout exists (this build
container); ``tests/test_oracle_vs_reference.py`` calls ``run_pin`` and skips
elsewhere.

Two legs:

1. free-running: the reference ``RocketLandingEnv`` (backend/envs/lander.py:188) and
   ``OracleEnv`` are driven by the same float32 action stream from the same injected
   initial states, resetting both from the same global-numpy seed on ``done``; every
   raw state word, observation, reward and flag must be IDENTICAL (``==`` on fp64).
2. teacher-forced edge cases: wild injected states (near the ground with every
   landing class, angle wraps at +-180, out of bounds, tipped over, empty tank, step
   999 -> max-step truncation) are stepped once in both and compared the same way.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from oracle.ref_loader import load_reference          # noqa: E402
from oracle import rocket_oracle as ro                # noqa: E402

REF_KEYS = ("x", "y", "vx", "vy", "ax", "ay", "angle", "angularVelocity", "mass", "fuelMass")


def _inject_reference(env, st: dict, current_step: int = 0):
    """set ``env.rocket.state`` then recompute the consistent previous
    state (backend/rocket/base.py:65)."""
    env.rocket.state = dict(st)
    env.rocket.previous_state = env.rocket.calculate_consistent_previous_state(
        env.rocket.state, env.rocket.dt)
    env.current_step = current_step


def _ref_state_vec(env):
    s = env.rocket.state
    return np.array([float(s[k]) for k in REF_KEYS])


def _oracle_state_vec(oenv):
    return np.array([getattr(oenv.state, k) for k in ro.STATE_WORDS])


def edge_case_states(rng: np.random.Generator, n: int):
    """Injected (state, action, current_step) triples that reach the branches random
    rollouts never visit."""
    cases = []
    for i in range(n):
        kind = i % 8
        st = {"x": rng.uniform(-1500, 1500), "y": rng.uniform(1800, 2200),
              "vx": rng.uniform(-25, 25), "vy": rng.uniform(-200, -150),
              "ax": 0.0, "ay": 0.0, "angle": rng.uniform(-15, 15),
              "angularVelocity": rng.uniform(-10, 10),
              "mass": rng.uniform(34000, 38000), "fuelMass": rng.uniform(370000, 410000)}
        step = 0
        if kind == 0:      # touchdown across all landing classes
            vy = -rng.uniform(0.5, 120)
            st.update(y=0.1 + rng.uniform(0.0, 1.0) * abs(vy) * 0.1, vy=vy,
                      vx=rng.uniform(-110, 110), angle=rng.uniform(-12, 12),
                      angularVelocity=rng.uniform(-3, 3))
        elif kind == 1:    # angle wrap through +-180
            sign = rng.choice([-1.0, 1.0])
            st.update(angle=sign * rng.uniform(170, 179.999),
                      angularVelocity=sign * rng.uniform(0, 300))
        elif kind == 2:    # out of bounds / very high
            if rng.random() < 0.5:
                st.update(x=rng.choice([-1.0, 1.0]) * rng.uniform(49990, 50010))
            else:
                st.update(y=rng.uniform(49990, 50010), vy=rng.uniform(-50, 200))
        elif kind == 3:    # tipped over, tumbling
            st.update(angle=rng.uniform(-180, 180), angularVelocity=rng.uniform(-400, 400))
        elif kind == 4:    # tank (almost) empty
            st.update(fuelMass=float(rng.choice([0.0, -1.0, 50.0, 170.0, 171.0])))
        elif kind == 5:    # last step before max-step truncation, climbing
            step = 999 - int(rng.integers(0, 2))
            st.update(vy=rng.uniform(-5, 30), y=rng.uniform(0.2, 1500))
        elif kind == 6:    # slow, near the ground, low throttle (free-fall term)
            st.update(y=rng.uniform(0.05, 30), vy=rng.uniform(-20, 5), vx=rng.uniform(-2, 2),
                      angle=rng.uniform(-11, 11), angularVelocity=rng.uniform(-0.2, 0.2))
        else:              # at rest on/below the ground, tiny rates (dead-band branches)
            st.update(y=rng.uniform(-5, 0.1), vy=rng.uniform(-1, 1) * 1e-5,
                      vx=rng.uniform(-1, 1) * 1e-5, angle=rng.uniform(-0.1, 0.1),
                      angularVelocity=rng.uniform(-0.1, 0.1))
        act = rng.uniform([-0.2, -1.3], [1.2, 1.3]).astype(np.float32)
        if i % 5 == 0:
            act[0] = np.float32(rng.choice([0.0, 0.1, 0.099999994, 1e-6, 1.0]))
        if i % 7 == 0:
            act[1] = np.float32(0.0)
        cases.append((st, act, step))
    return cases


def run_pin(n_free_envs: int = 6, free_steps: int = 400, n_edge: int = 4000,
            verbose: bool = False) -> dict:
    ref = load_reference()
    consts = ro.EnvConstants.from_yaml(os.path.join(ref.root, "config.yaml"))
    # the repo's own config.yaml must describe the same environment
    own = ro.EnvConstants.from_yaml()
    assert own == consts, "config.yaml in this repo disagrees with the reference's constants"

    checked = 0
    # ---- leg 1: free-running with seeded resets -------------------------------------
    for e in range(n_free_envs):
        renv = ref.RocketLandingEnv()
        oenv = ro.OracleEnv(consts)
        np.random.seed(1000 + e)
        robs, _ = renv.reset()
        np.random.seed(1000 + e)
        oobs, _ = oenv.reset()
        assert np.array_equal(robs, oobs)
        assert np.array_equal(_ref_state_vec(renv), _oracle_state_vec(oenv))
        rng = np.random.default_rng(e)
        for t in range(free_steps):
            a = rng.uniform([0, -1], [1, 1]).astype(np.float32)
            robs, rr, rterm, rtrunc, _ = renv.step(a)
            oobs, orr, oterm, otrunc, _ = oenv.step(a)
            assert np.array_equal(_ref_state_vec(renv), _oracle_state_vec(oenv)), (e, t)
            assert np.array_equal(robs, oobs), (e, t)
            assert rr == orr and rterm == oterm and rtrunc == otrunc, (e, t, rr, orr)
            p = renv.rocket.previous_state
            assert (p["x"], p["y"], p["angle"]) == (oenv.state.prev_x, oenv.state.prev_y,
                                                    oenv.state.prev_angle), (e, t)
            checked += 1
            if rterm or rtrunc:
                seed = 50000 + 1000 * e + t
                np.random.seed(seed)
                robs, _ = renv.reset()
                np.random.seed(seed)
                oobs, _ = oenv.reset()
                assert np.array_equal(robs, oobs)

    # ---- leg 2: teacher-forced edge cases ----------------------------------------------
    renv = ref.RocketLandingEnv()
    oenv = ro.OracleEnv(consts)
    branch = {"terminated": 0, "truncated": 0, "safe": 0, "good": 0, "ok": 0, "unsafe": 0,
              "wrap": 0, "tip": 0}
    for st, act, step in edge_case_states(np.random.default_rng(7), n_edge):
        _inject_reference(renv, st, step)
        oenv.inject(st["x"], st["y"], st["vx"], st["vy"], st["angle"], st["angularVelocity"],
                    st["mass"], st["fuelMass"], current_step=step)
        for _ in range(2):   # second step exercises the carried (normalised) previous angle
            before_angle = oenv.state.angle
            robs, rr, rterm, rtrunc, _ = renv.step(act)
            oobs, orr, oterm, otrunc, info = oenv.step(act)
            assert np.array_equal(_ref_state_vec(renv), _oracle_state_vec(oenv)), st
            assert np.array_equal(robs, oobs), st
            assert rr == orr and rterm == oterm and rtrunc == otrunc, (st, rr, orr)
            checked += 1
            branch["terminated"] += rterm
            branch["truncated"] += rtrunc
            if rterm:
                msg = ref.evaluate_landing(renv.rocket.get_state(), renv.config)["landing_message"]
                assert msg == ro.LANDING_NAMES[info["landing"]]
                branch[msg] += 1
            branch["wrap"] += abs(oenv.state.angle - before_angle) > 180
            branch["tip"] += abs(oenv.state.angle) > 90
    for k, v in branch.items():
        assert v > 0, f"edge-case generator never reached branch '{k}'"
    if verbose:
        print(f"pin OK: {checked} steps identical; branches {branch}")
    return {"steps_checked": checked, **branch}


if __name__ == "__main__":
    run_pin(verbose=True)
