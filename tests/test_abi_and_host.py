"""CPU tier: the C-ABI library loads and exports every symbol the header declares (no
compute calls without a GPU), and the host-side logic (config loader, spaces, state-format
helpers, sharding, statistics)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import rocket_landing_rl_b200 as pkg
from rocket_landing_rl_b200 import _lib, layout
from rocket_landing_rl_b200.config import Config, RlParams, load_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rocket_landing_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rl_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert lib.rl_abi_version() == _lib.ABI_VERSION == 2


def test_header_constants_match_python_layout():
    text = open(HEADER).read()
    defs = dict(re.findall(r"#define\s+(RL_[A-Z_]+)\s+(-?\(?-?[0-9xA-Fa-fu]+\)?)", text))
    val = lambda k: int(defs[k].strip("()").rstrip("u"), 0)   # noqa: E731
    assert val("RL_STATE_WORDS") == len(layout.STATE_WORDS) == 10
    for name, idx in (("RL_W_X", layout.W_X), ("RL_W_Y", layout.W_Y), ("RL_W_ANGLE", layout.W_ANGLE),
                      ("RL_W_SX", layout.W_SX), ("RL_W_SY", layout.W_SY), ("RL_W_SANGLE", layout.W_SANGLE),
                      ("RL_W_DRY", layout.W_DRY), ("RL_W_FUEL", layout.W_FUEL), ("RL_W_META", layout.W_META),
                      ("RL_W_EPRET", layout.W_EPRET)):
        assert val(name) == idx
    assert val("RL_META_STEP_MASK") == layout.META_STEP_MASK
    assert val("RL_META_FRESH") == layout.META_FRESH
    assert val("RL_META_DLO_SHIFT") == layout.META_DLO_SHIFT
    assert val("RL_META_WRAP_SHIFT") == layout.META_WRAP_SHIFT
    assert "RL_ANGLE_UNITS_PER_DEG (4294967296.0 / 360.0)" in text and layout.ANGLE_UNITS_PER_DEG == 4294967296.0 / 360.0
    assert val("RL_DELTA_UNITS_PER_POS_UNIT") == layout.DELTA_UNITS_PER_POS_UNIT == 256
    # the fixed-point fuel unit: library and Python helper derive the same power of two
    lib = _lib.load()
    p = load_params()
    assert lib.rl_fuel_unit_for(C.byref(p)) == layout.fuel_unit_for(p.reset_high[9]) == layout.DEFAULT_FUEL_UNIT
    p = load_params()
    assert lib.rl_pos_unit_for(C.byref(p)) == layout.pos_unit_for(p) == layout.DEFAULT_POS_UNIT == 2.0 ** -15
    for world in (10.0, 50412.0, 50413.0, 2.0e6):           # 1.3 x 50 412.3 = 65 536
        p.max_horizontal_position = world
        assert lib.rl_pos_unit_for(C.byref(p)) == layout.pos_unit_for(p)
    p = load_params()
    for hi in (0.5, 1.0, 3.0, 524287.0, 524288.0, 1.0e9):
        p.reset_high[9] = hi
        u = lib.rl_fuel_unit_for(C.byref(p))
        assert u == layout.fuel_unit_for(hi) and hi / u < 2.0 ** 31 <= 4.0 * max(hi, 0.5) / u
    from rocket_landing_rl_b200.stats import SLOTS, STATS_WORDS
    assert val("RL_STATS_WORDS") == STATS_WORDS
    for i, slot in enumerate(SLOTS):
        assert val("RL_S_" + slot.upper()) == i


def test_params_default_equals_yaml():
    lib = _lib.load()
    d = RlParams()
    lib.rl_params_default(C.byref(d))
    assert bytes(d) == bytes(load_params())
    assert lib.rl_state_bytes(1) == 40 * 256 and lib.rl_state_bytes(257) == 40 * 512
    assert lib.rl_state_bytes(0) == 0


def test_create_fails_loudly_without_gpu_or_with_bad_args():
    import torch
    lib = _lib.load()
    h = C.c_void_p()
    p = load_params()
    assert lib.rl_create(C.byref(p), 0, 0, 0, 0, None, C.byref(h)) == -1
    assert b"n_envs" in lib.rl_last_error()
    bad = load_params()
    bad.time_step = 0.0
    assert lib.rl_create(C.byref(bad), 16, 0, 0, 0, None, C.byref(h)) == -1
    assert b"time_step" in lib.rl_last_error()
    if not torch.cuda.is_available():
        assert lib.rl_create(C.byref(p), 16, 0, 0, 0, None, C.byref(h)) == -2     # RL_E_CUDA
        assert lib.rl_last_error()
        with pytest.raises(_lib.RocketLibraryError):
            pkg.RocketLandingVecEnv(16)
        with pytest.raises(_lib.RocketLibraryError):
            pkg.RocketLandingSB3VecEnv(16)
    assert lib.rl_step(None, None, None, None, None, None, None, None, None, None, None, 1, 1, None) == -1
    # the rollout neighbours validate their arguments before they touch the device
    assert lib.rl_episode_flags(None, None, None, None, None, 16, None) == -1 and b"null" in lib.rl_rollout_last_error()
    assert lib.rl_gae(None, None, None, None, None, None, None, 1, 16, 0.99, 0.95, None) == -1
    assert lib.rl_mlp_forward(None, None, None, 16, 1, None) == -1 and b"null" in lib.rl_policy_last_error()
    assert lib.rl_policy_value_sample(None, None, None, None, 0, 0, 0, None, None, None, None, None, None, None, 16, None) == -1
    assert b"rl_policy_value_sample" in lib.rl_policy_last_error()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg_dir = os.path.join(ROOT, "rocket-landing-rl_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "liboracle" not in src and "hostemu.cpp" not in src.replace("tests/hostemu", ""), f


# ---- config loader: strictness of backend/config.py:21-37 ---------------------------------------
def test_config_strict_lookup(tmp_path):
    cfg = Config()
    assert cfg.get("simulation.time_step") == 0.1
    assert cfg.get("rocket.position_limits.x") == [-1500.0, 1500.0]
    with pytest.raises(KeyError):
        cfg.get("rocket.nope")
    with pytest.raises(KeyError):
        cfg.get("simulation.time_step.deeper")
    y = tmp_path / "c.yaml"
    y.write_text("a:\n  b:\n")
    with pytest.raises(ValueError):
        Config(str(y)).get("a.b")
    with pytest.raises(FileNotFoundError):
        Config(str(tmp_path / "missing.yaml"))


def test_load_params_reads_every_constant():
    p = load_params()
    assert (p.time_step, p.gravity, p.thrust_power, p.cold_gas_thrust_power) == (0.1, -9.81, 1e7, 7e4)
    assert list(p.obs_low) == [-1500.0, 0.0, -25.0, -200.0, -5.0, -5.0, -15.0, -10.0]
    assert list(p.obs_high) == [1500.0, 2200.0, 25.0, -150.0, 5.0, 5.0, 15.0, 10.0]
    assert list(p.land_ok) == [100.0, 100.0, 10.0] and p.max_episode_steps == 1000


def test_box_space():
    b = pkg.Box(low=np.array([0.0, -1.0], np.float32), high=np.array([1.0, 1.0], np.float32))
    assert b.shape == (2,) and b.dtype == np.float32
    s = b.sample()
    assert s.dtype == np.float32 and b.contains(s) and not b.contains(np.array([2.0, 0.0], np.float32))


# ---- public state format -------------------------------------------------------------------------
def test_fresh_state_round_trip():
    soa = layout.fresh_state_soa([1.0, 2.0], [2000.0, 1900.0], [3.0, -3.0], [-160.0, -170.0],
                                 [5.0, -5.0], [1.0, -1.0], [35000.0, 36000.0], [4e5, 3.9e5], steps=[0, 7])
    v = layout.reference_view(soa, 0.1)
    assert v["fresh"].all() and v["steps"].tolist() == [0, 7] and v["wraps"].tolist() == [0, 0]
    np.testing.assert_allclose(v["vy"], [-160.0, -170.0])
    np.testing.assert_allclose(v["ang_vel"], [1.0, -1.0])
    np.testing.assert_allclose(v["angle"], [5.0, -5.0], atol=1e-7)
    np.testing.assert_allclose(v["fuel"], [4e5, 3.9e5], atol=1e-4)


def test_angle_delta_keeps_32_bits():
    """The carried angle delta is fp32 + 8 low-order bits in the meta word (layout.split / join,
    csrc/rl_device.cuh dtheta_split / dtheta_join): relative error <= 2^-32 (2^-31 for the 0.4 % of
    values whose low part saturates at +127)."""
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.uniform(-30, 30, 5000), rng.uniform(-1e-3, 1e-3, 5000), [0.0, 1.0, -1.0, 2.0 ** -140]])
    hi, lo = layout.split_angle_delta(v)
    assert lo.min() >= -128 and lo.max() <= 127
    back = layout.join_angle_delta(hi, lo)
    rel = np.abs(back - v) / np.maximum(np.abs(v), 1e-300)
    assert rel.max() <= 2.0 ** -31 and np.mean(rel > 2.0 ** -32) < 0.01
    assert np.abs(hi.astype(np.float64) - v).max() > 100 * np.abs(back - v).max()


def test_mid_episode_state_from_reference_dicts(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_edge_cases.npz"))
    state, prev = g["state"][0], g["prev"][0]
    soa = layout.state_soa_from_reference(state, prev, steps=3, dt=0.1)
    v = layout.reference_view(soa, 0.1)
    assert not v["fresh"].any() and (v["steps"] == 3).all()
    np.testing.assert_allclose(v["vx"], state[:, 2], rtol=2e-6, atol=1e-5)
    np.testing.assert_allclose(v["ang_vel"], state[:, 7], rtol=2e-6, atol=1e-5)
    # wrap count: (angle - prev_angle) = pre-wrap delta - 360 k
    delta = state[:, 6] - prev[:, 2]
    np.testing.assert_allclose(state[:, 7] * 0.1 - 360.0 * v["wraps"], delta, atol=1e-6)
    assert (v["wraps"] != 0).sum() > 50


# ---- sharding + statistics -----------------------------------------------------------------------
def test_shard_range_partitions_exactly():
    for total, world in ((64, 8), (67_108_864, 8), (10, 3), (7, 7), (1000, 1)):
        spans = [pkg.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (s0, c0), (s1, _) in zip(spans, spans[1:]):
            assert s0 + c0 == s1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        pkg.shard_range(3, 0, 4)
    with pytest.raises(ValueError):
        pkg.shard_range(8, 8, 8)


def test_episode_stats_vector():
    vec = np.zeros(16)
    vec[:13] = [10, 50.0, 1200, 1, 2, 3, 4, 0, 0, 5, 0, 1200, 60.0]
    st = pkg.EpisodeStats.from_vector(vec)
    assert st.mean_return == 5.0 and st.mean_length == 120.0
    assert st.landing_success_rate == pytest.approx(0.6)
    assert np.isnan(pkg.EpisodeStats().mean_return)
