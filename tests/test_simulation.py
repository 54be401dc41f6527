"""Simulation mode + telemetry: the oracle against the
reference's own RocketControls / BinaryProtocol and its committed telemetry stream; the device
math (host build on CPU, CUDA on the GPU) against the reference's stream."""
import ctypes as C
import os

import numpy as np
import pytest

import parity
from oracle import simulation_oracle as so
from oracle.ref_loader import reference_available
from rocket_landing_rl_b200 import layout

FIELDS = ("x", "y", "vx", "vy", "ax", "ay", "angle", "angularVelocity", "angularAcceleration", "mass",
          "fuelMass", "reward", "throttle", "coldGas", "landing_code", "is_active")
# tolerance per record word: the state scales of tests/parity.py, alpha ~ 10 deg/s^2
SCALE = np.array([1500, 2000, 25, 200, 10, 10, 15, 10, 10, 36000, 390000, 0, 1, 1, 1, 1], float)


@pytest.fixture(scope="module")
def golden(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "ref_telemetry.npz")))


def test_oracle_reproduces_the_reference_telemetry_stream_byte_for_byte(golden):
    sim = so.SimulationOracle(golden["init_state"].shape[0])
    sim.inject(golden["init_state"])
    for t in range(golden["actions"].shape[0]):
        msg = sim.step(golden["actions"][t])
        assert msg == golden["messages"][t].tobytes(), f"tick {t}"
    rec = so.decode(msg)
    assert set(rec[:, 14].tolist()) == {1.0, 4.0} and rec[:, 15].sum() == 0      # safe and unsafe landings, all done
    assert np.isnan(rec[:, 11]).all()


@pytest.mark.skipif(not reference_available(), reason="reference checkout not present (GPU box)")
def test_oracle_controls_and_protocol_against_the_live_reference():
    import tempfile
    from oracle.ref_loader import load_reference
    from oracle.pin_against_reference import REF_KEYS
    load_reference()
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        from backend.rocket.controls import RocketControls
        from backend.protocol import BinaryProtocol
        rng = np.random.default_rng(5)
        for trial in range(6):
            rc, oc = RocketControls(), so.ControlsOracle()
            s = rng.uniform([-1500, 40, -25, -60, 0, 0, -15, -10, 34000, 370000],
                            [1500, 400, 25, -20, 0, 0, 15, 10, 38000, 410000])
            rc.rocket.state = {k: float(v) for k, v in zip(REF_KEYS, s)}
            rc.rocket.previous_state = rc.rocket.calculate_consistent_previous_state(rc.rocket.state, rc.rocket.dt)
            oc.inject(s[0], s[1], s[2], s[3], s[6], s[7], s[8], s[9])
            for t in range(120):
                a = rng.uniform([-0.3, -1.4], [1.3, 1.4]).astype(np.float32)
                rs, rr, rd = rc.step(a)
                os_, orr, od, cls = oc.step(a)
                assert (rr, rd) == (orr, od), (trial, t)
                assert [rs[k] for k in ("x", "y", "vx", "vy", "angle", "angularVelocity", "fuelMass")] == \
                       [os_.x, os_.y, os_.vx, os_.vy, os_.angle, os_.ang_vel, os_.fuel]
                want = BinaryProtocol.encode_rocket_state(rs, rr, {"throttle": float(a[0]), "coldGas": float(a[1])},
                                                          "safe" if cls == 1 else None)
                assert want == so.encode_record(os_, orr, (float(a[0]), float(a[1])), 1 if cls == 1 else 0)
            assert rc.step(a)[1:] == (0.0, True) or not rd           # frozen after touchdown
        assert BinaryProtocol.encode_rocket_state(None, None, {}, "unsafe") == so.encode_record(None, None, (0, 0), 4)
        assert BinaryProtocol.encode_telemetry_header() == bytes([so.MSG_TELEMETRY])
    finally:
        os.chdir(cwd)


def _compare_stream(step_fn, golden):
    """step_fn(actions) -> [N,16] records; compare every tick with the reference's message."""
    worst = np.zeros(16)
    mismatched_ticks = 0
    for t in range(golden["actions"].shape[0]):
        got = step_fn(golden["actions"][t])
        want = so.decode(golden["messages"][t].tobytes())
        # discrete words: active flag, landing code, echoed action -- exact unless the touchdown tick differs
        same_tick = np.array_equal(got[:, 15], want[:, 15])
        if not same_tick:
            mismatched_ticks += 1
            continue
        np.testing.assert_array_equal(got[:, 14], want[:, 14], err_msg=f"landing code, tick {t}")
        np.testing.assert_array_equal(got[:, 12:14], want[:, 12:14], err_msg=f"echoed action, tick {t}")
        assert np.array_equal(np.isnan(got[:, 11]), np.isnan(want[:, 11]))
        act = want[:, 15] == 1.0
        tol = parity.RTOL * np.maximum(np.abs(want[act, :11].astype(np.float64)), SCALE[:11])
        err = np.abs(got[act, :11].astype(np.float64) - want[act, :11]) / tol
        if err.size:
            worst[:11] = np.maximum(worst[:11], err.max(axis=0))
        rtol = np.maximum(parity.FREE_REWARD_ATOL, parity.FREE_REWARD_RTOL * np.abs(want[act, 11]))
        assert (np.abs(got[act, 11] - want[act, 11]) <= rtol).all(), f"reward, tick {t}"
    assert (worst <= 1.0).all(), dict(zip(FIELDS, worst.round(3)))
    assert mismatched_ticks == 0, "a touchdown tick differed (outside what the guard band allows here)"
    return worst


def test_host_build_of_the_sim_mode_matches_the_reference_stream(golden):
    n = golden["init_state"].shape[0]
    st = parity.HostEmuStepper(n)
    st.inject_fresh(golden["init_state"])
    st.L.emu_sim_step.argtypes = [C.c_void_p] * 2 + [C.c_int64] + [C.c_void_p] * 2

    def step(a):
        rec = np.zeros((n, 16), np.float32)
        a = np.ascontiguousarray(a, np.float32)
        assert st.L.emu_sim_step(C.addressof(st.params), st.soa.ctypes.data, n, a.ctypes.data, rec.ctypes.data) == 0
        return rec

    worst = _compare_stream(step, golden)
    print("sim mode (host build) worst err / tol:", dict(zip(FIELDS[:11], worst[:11].round(3))))


def _edge_states():
    n = 4
    init = np.zeros((n, 10))
    init[:, 0] = [0.0, 65000.0, -65000.0, 100.0]                 # x
    init[:, 1] = [64000.0, 1000.0, 1000.0, 2000.0]               # y
    init[:, 2] = [0.0, 240.0, -240.0, 0.0]                       # vx
    init[:, 3] = [240.0, 0.0, 0.0, -10.0]                        # vy
    init[:, 8], init[:, 9] = 36000.0, 400000.0
    return init


def _check_edge(xs, ys):
    xs, ys = np.array(xs), np.array(ys)
    assert (np.diff(ys[:, 0]) >= 0).all() and ys[-1, 0] == pytest.approx(65536.0, abs=1e-2)     # climbs, then sits at the top
    assert (np.diff(xs[:, 1]) >= 0).all() and xs[-1, 1] == pytest.approx(65536.0, abs=1e-2)
    assert (np.diff(xs[:, 2]) <= 0).all() and xs[-1, 2] == pytest.approx(-65536.0, abs=1e-2)
    assert abs(xs[-1, 3] - 100.0) < 1.0 and ys[:, 3].max() < 60000.0                          # an ordinary rocket is untouched


def test_sim_mode_rocket_stops_at_the_edge_of_the_position_grid():
    """The simulation mode has no bounds check (controls.py:89 ignores truncation), so a rocket under
    full throttle eventually leaves the fixed-point position grid (+-65 536 m for the reference's
    world): it must stay at the edge, not wrap around to the other side."""
    init = _edge_states()
    n = init.shape[0]
    st = parity.HostEmuStepper(n)
    st.L.emu_sim_step.argtypes = [C.c_void_p] * 2 + [C.c_int64] + [C.c_void_p] * 2
    st.inject_fresh(init)
    a = np.tile(np.array([[1.0, 0.0]], np.float32), (n, 1))
    rec = np.zeros((n, 16), np.float32)
    xs, ys = [], []
    for t in range(400):
        assert st.L.emu_sim_step(C.addressof(st.params), st.soa.ctypes.data, n, a.ctypes.data, rec.ctypes.data) == 0
        xs.append(rec[:, 0].copy())
        ys.append(rec[:, 1].copy())
    _check_edge(xs, ys)


@pytest.mark.gpu
def test_cuda_sim_mode_rocket_stops_at_the_edge_of_the_position_grid():
    from rocket_landing_rl_b200 import RocketSimulation
    s = _edge_states()
    n = s.shape[0]
    sim = RocketSimulation(n)
    sim.set_state(layout.fresh_state_soa(s[:, 0], s[:, 1], s[:, 2], s[:, 3], s[:, 6], s[:, 7], s[:, 8], s[:, 9]))
    a = np.tile(np.array([[1.0, 0.0]], np.float32), (n, 1))
    xs, ys = [], []
    for t in range(400):
        rec = RocketSimulation.decode(sim.step(a))
        xs.append(rec[:, 0].copy())
        ys.append(rec[:, 1].copy())
    _check_edge(xs, ys)
    sim.close()


@pytest.mark.gpu
def test_cuda_sim_mode_matches_the_reference_stream(golden):
    from rocket_landing_rl_b200 import RocketSimulation
    n = golden["init_state"].shape[0]
    sim = RocketSimulation(n)
    s = golden["init_state"]
    sim.set_state(layout.fresh_state_soa(s[:, 0], s[:, 1], s[:, 2], s[:, 3], s[:, 6], s[:, 7], s[:, 8], s[:, 9]))
    msgs = []

    def step(a):
        m = sim.step(a)
        msgs.append(m)
        assert m[0] == 1 and len(m) == 1 + 64 * n
        return RocketSimulation.decode(m)

    worst = _compare_stream(step, golden)
    print("sim mode (CUDA) worst err / tol:", dict(zip(FIELDS[:11], worst[:11].round(3))))
    # reset re-arms every rocket
    obs = sim.reset()
    assert obs.shape == (n, 8)
    rec = RocketSimulation.decode(sim.step(np.zeros((n, 2), np.float32)))
    assert (rec[:, 15] == 1.0).all() and (rec[:, 14] == 0.0).all()
    sim.close()
