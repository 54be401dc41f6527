"""The oracle against (a) the live reference where it exists, (b) the committed outputs of
the reference, (c) the reference's own physics unit tests restated.  CPU only."""
import os

import numpy as np
import pytest

from oracle import c_oracle, rocket_oracle as ro
from oracle.ref_loader import reference_available


@pytest.fixture(scope="module")
def consts():
    return ro.EnvConstants.from_yaml()


@pytest.mark.skipif(not reference_available(), reason="reference checkout not present (GPU box)")
def test_python_oracle_is_bit_identical_to_live_reference():
    from oracle.pin_against_reference import run_pin
    out = run_pin(n_free_envs=3, free_steps=300, n_edge=1600)
    assert out["steps_checked"] >= 3 * 300 + 2 * 1600


def _replay_python_oracle(g, e, consts):
    env = ro.OracleEnv(consts)
    s0 = g["init_state"][e] if g["init_state"].ndim == 2 else g["init_state"]
    env.inject(s0[0], s0[1], s0[2], s0[3], s0[6], s0[7], s0[8], s0[9], ax=s0[4], ay=s0[5])
    T = g["actions"].shape[0]
    pick = (lambda k, t: g[k][t, e]) if g["actions"].ndim == 3 else (lambda k, t: g[k][t])
    for t in range(T):
        obs, r, term, trunc, info = env.step(pick("actions", t))
        st = np.array([getattr(env.state, k) for k in ro.STATE_WORDS])
        ref = pick("state", t)
        assert term == pick("terminated", t) and trunc == pick("truncated", t), (e, t)
        assert info["landing"] == pick("landing", t), (e, t)
        np.testing.assert_allclose(st, ref, rtol=1e-12, atol=1e-12, err_msg=f"state e={e} t={t}")
        assert abs(r - pick("reward", t)) <= 1e-9 * max(1.0, abs(r))
        done = term or trunc
        ref_obs = pick("terminal_obs", t) if done else pick("obs", t)
        np.testing.assert_allclose(obs, ref_obs, rtol=1e-6, atol=1e-6)
        if done:
            s = pick("reset_state", t)
            reset_obs = env.inject(s[0], s[1], s[2], s[3], s[6], s[7], s[8], s[9], ax=s[4], ay=s[5])
            np.testing.assert_allclose(reset_obs, pick("obs", t), rtol=1e-6, atol=1e-6)


def test_python_oracle_reproduces_reference_rollout_single(golden_dir, consts):
    g = np.load(os.path.join(golden_dir, "ref_rollout_single.npz"))
    assert int((g["terminated"] | g["truncated"]).sum()) >= 5
    _replay_python_oracle(g, 0, consts)


def test_python_oracle_reproduces_reference_rollout_batch(golden_dir, consts):
    g = np.load(os.path.join(golden_dir, "ref_rollout_batch8.npz"))
    for e in (0, 5):
        _replay_python_oracle(g, e, consts)


def test_c_oracle_reproduces_reference_rollouts(golden_dir, consts):
    g = np.load(os.path.join(golden_dir, "ref_rollout_batch8.npz"))
    T, E = g["actions"].shape[:2]
    b = c_oracle.BatchOracle(E, consts)
    b.inject(g["init_state"])
    for t in range(T):
        obs, rew, term, trunc, land = b.step(g["actions"][t])
        assert np.array_equal(term, g["terminated"][t]) and np.array_equal(trunc, g["truncated"][t]), t
        assert np.array_equal(land, g["landing"][t]), t
        np.testing.assert_allclose(b.state(), g["state"][t], rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(rew, g["reward"][t], rtol=1e-9, atol=1e-9)
        done = term | trunc
        ref_obs = np.where(done[:, None], g["terminal_obs"][t], g["obs"][t])
        np.testing.assert_allclose(obs, ref_obs, rtol=1e-6, atol=1e-6)
        if done.any():
            b.inject(g["reset_state"][t], mask=done)


def test_c_oracle_reproduces_reference_edge_cases(golden_dir, consts):
    g = np.load(os.path.join(golden_dir, "ref_edge_cases.npz"))
    n = g["init_state"].shape[0]
    b = c_oracle.BatchOracle(n, consts)
    b.inject(g["init_state"], current_step=g["init_step"])
    seen = np.zeros(5, int)
    for k in range(2):
        obs, rew, term, trunc, land = b.step(g["actions"])
        assert np.array_equal(term, g["terminated"][k]) and np.array_equal(trunc, g["truncated"][k])
        assert np.array_equal(land, g["landing"][k])
        np.testing.assert_allclose(b.state(), g["state"][k], rtol=1e-11, atol=1e-9)
        np.testing.assert_allclose(rew, g["reward"][k], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(obs, g["obs"][k], rtol=1e-6, atol=1e-6)
        seen += np.bincount(land, minlength=5)
    assert (seen[1:] > 0).all(), "every landing class must be exercised"


# ---- the reference's tests/test_physics.py:15-147, restated against the oracle ---------------
def test_gravity_force(consts):                              # test_physics.py:15
    fx, fy = ro.net_force(consts, 1000.0, 0.0, 0.0, 0.0, 0.0)
    assert fx == pytest.approx(0.0) and fy == pytest.approx(1000.0 * consts.gravity)


def test_thrust_no_throttle(consts):                         # :23
    fx, fy = ro.net_force(consts, 1.0, 0.0, 0.0, 0.0, 0.0)
    assert (fx, fy) == pytest.approx((0.0, consts.gravity))


def test_thrust_full_vertical_and_angled(consts):            # :34, :46
    m = 2.0
    fx, fy = ro.net_force(consts, m, 1.0, 0.0, 0.0, 0.0)
    assert fx == pytest.approx(0.0) and fy == pytest.approx(consts.thrust_power + m * consts.gravity)
    fx, fy = ro.net_force(consts, m, 1.0, 45.0, 0.0, 0.0)
    assert fx == pytest.approx(consts.thrust_power * np.sin(np.deg2rad(45)))
    assert fy == pytest.approx(consts.thrust_power * np.cos(np.deg2rad(45)) + m * consts.gravity)


def test_drag_zero_and_opposing(consts):                     # :59, :67
    fx, fy = ro.net_force(consts, 1.0, 0.0, 0.0, 0.0, 0.0)
    assert fx == 0.0
    fx, fy = ro.net_force(consts, 1.0, 0.0, 0.0, 10.0, -10.0)
    dx, dy = fx, fy - consts.gravity
    assert np.hypot(dx, dy) > 0 and dx * 10.0 + dy * -10.0 < 0


def test_acceleration_is_force_over_mass(consts):            # :88
    assert ro.linear_acceleration(10.0, -20.0, 2.0) == pytest.approx((5.0, -10.0))
    assert ro.linear_acceleration(10.0, -20.0, 0.0) == (0.0, 0.0)


def test_verlet_step_gravity_only(consts):                   # :97
    s = ro.RocketState(x=0.0, y=100.0, dry_mass=1.0, fuel=0.0, prev_x=0.0, prev_y=100.0)
    ro.apply_action(consts, s, 0.0, 0.0)
    assert s.x == pytest.approx(0.0)
    assert s.y == pytest.approx(2 * 100.0 - 100.0 + consts.gravity * consts.dt ** 2)
    assert s.vy == pytest.approx((s.y - 100.0) / consts.dt)


def test_fuel_consumption(consts):                           # :142
    s = ro.RocketState(y=100.0, dry_mass=1000.0, fuel=500.0, prev_y=100.0)
    ro.apply_action(consts, s, 0.5, 0.0)
    assert 500.0 - s.fuel == pytest.approx(0.5 * consts.fuel_rate * consts.dt)


def test_normalize_angle_range():
    for a, want in ((0.0, 0.0), (180.0, -180.0), (-180.0, -180.0), (190.0, -170.0), (-190.0, 170.0),
                    (540.0, -180.0), (359.0, -1.0)):
        assert ro.normalize_angle_180(a) == pytest.approx(want)


def test_c_oracle_multithreaded_rollout_runs(consts):
    out = c_oracle.rollout_random(32, 300, 2, seed=1, constants=consts)
    assert out["env_steps"] == 2 * 32 * 300 and out["episodes"] > 50
    assert out["landings"][4] > 0
