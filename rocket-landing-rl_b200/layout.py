"""Host-side helpers for the public state format of ``rl_set_state`` / ``rl_get_state``
(32-bit words ``[10][N]`` in an int32 array -- ``f32`` words hold their IEEE-754 bits -- words
``RL_W_*`` of include/rocket_landing_b200.h).

The kernel does not store the reference's ``Rocket.state`` dict verbatim: it carries the
Verlet DELTAS ``x(t) - x(t-dt)`` (from which ``v = delta / dt``, engine.py:193-194) instead of
``previous_state`` (base.py:171), and a FRESH flag instead of a bootstrapped previous state
(base.py:65): a FRESH rocket's deltas are ``v dt``, and its first step subtracts the bootstrap
term ``a0 dt^2 / 2`` -- see DESIGN.md "fp32 formulation".  The words that integrate over an episode are
32-bit FIXED POINT (angle, position deltas, fuel) or carry 8 extra mantissa bits (angle delta),
DESIGN.md "Long-horizon accuracy".  These helpers convert in both directions, in float64, so
nothing is lost before the final rounding to the device word.
"""
from __future__ import annotations

import math

import numpy as np

STATE_WORDS = ("x", "y", "angle", "sx", "sy", "sangle", "dry_mass", "fuel", "meta", "ep_return")
W_X, W_Y, W_ANGLE, W_SX, W_SY, W_SANGLE, W_DRY, W_FUEL, W_META, W_EPRET = range(10)
META_STEP_MASK = 0x0000FFFF
META_DLO_SHIFT = 16
META_WRAP_SHIFT = 24
META_FRESH = 0x80000000
ANGLE_UNITS_PER_DEG = 4294967296.0 / 360.0       # int32 range == [-180, 180)
DELTA_UNITS_PER_POS_UNIT = 256                   # the position deltas count 1/256 of the position unit
DEFAULT_POS_UNIT = 1.0 / 32768.0                 # m; reference config.yaml (+-50 km world): deltas in 2^-23 m
DEFAULT_FUEL_UNIT = 1.0 / 4096.0                 # kg; reference config.yaml (fuel <= 410 000 kg)


def fuel_unit_for(max_fuel: float) -> float:
    """kg per unit of the signed fixed-point fuel word (``rl_fuel_unit_for``): 2^(k-31) with 2^k
    the first power of two above the largest fuel load of the reset distribution."""
    k = int(math.floor(math.log2(max_fuel))) + 1 if max_fuel >= 1.0 else 0
    return math.ldexp(1.0, k - 31)


def pos_unit_for(params) -> float:
    """m per unit of the fixed-point positions (``rl_pos_unit_for``) of an ``RlParams``."""
    m = max(abs(params.max_horizontal_position), abs(params.max_altitude),
            *(abs(v) for k in range(2) for v in (params.reset_low[k], params.reset_high[k]))) * 1.3
    k = int(math.floor(math.log2(m))) + 1 if m >= 1.0 else 0
    return math.ldexp(1.0, k - 31)


def _pos_units(metres, pos_unit: float) -> np.ndarray:
    u = np.rint(np.asarray(metres, np.float64) / pos_unit)
    return np.clip(u, -2147483648, 2147483647).astype(np.int32)


def _bits(a) -> np.ndarray:
    """float values -> the int32 words holding their float32 bits; integer arrays pass through."""
    a = np.asarray(a)
    if a.dtype.kind == "f":
        return np.ascontiguousarray(a, np.float32).view(np.int32)
    return np.ascontiguousarray(a).astype(np.int64).astype(np.uint32, casting="unsafe").view(np.int32) \
        if a.dtype != np.int32 else a


def as_words(soa) -> np.ndarray:
    """Any 32-bit [10, N] array (int32 / uint32 / float32 bits) -> int32 words."""
    soa = np.asarray(soa)
    if soa.dtype.itemsize != 4:
        raise TypeError("state words must be a 32-bit array")
    return np.ascontiguousarray(soa).view(np.int32)


def f32_word(words: np.ndarray, w: int) -> np.ndarray:
    """Word ``w`` of an int32 state array as the float32 values it holds."""
    return words[w].view(np.float32)


def _meta_bits(steps, wraps, fresh, dlo=0) -> np.ndarray:
    steps = np.asarray(steps, np.int64) & META_STEP_MASK
    wraps = (np.asarray(wraps, np.int64) & 7) << META_WRAP_SHIFT
    dlo = (np.asarray(dlo, np.int64) & 0xFF) << META_DLO_SHIFT
    bits = (steps | wraps | dlo | (META_FRESH if fresh else 0)).astype(np.uint32)
    return bits.view(np.int32)


def _angle_units(angle_deg: np.ndarray) -> np.ndarray:
    """degrees in [-180, 180) -> int32 units (wraps like the kernel's integer add)."""
    u = np.rint(np.asarray(angle_deg, np.float64) * ANGLE_UNITS_PER_DEG).astype(np.int64)
    return (u & 0xFFFFFFFF).astype(np.uint32).view(np.int32)


def _fuel_units(fuel_kg: np.ndarray, fuel_unit: float) -> np.ndarray:
    u = np.rint(np.asarray(fuel_kg, np.float64) / fuel_unit)
    return np.clip(u, -2147483648, 2147483647).astype(np.int32)


def split_angle_delta(dang: np.ndarray):
    """float64 angle delta -> (float32 high part, int low part in [-128, 127]); the low part
    counts units of ulp(high) / 256 (``dtheta_split`` in csrc/rl_device.cuh)."""
    dang = np.asarray(dang, np.float64)
    hi = dang.astype(np.float32)
    e = ((hi.view(np.uint32) >> 23) & 0xFF).astype(np.int64)
    unit = np.ldexp(1.0, (e - 158).astype(np.int32))
    lo = np.clip(np.rint((dang - hi.astype(np.float64)) / unit), -128, 127).astype(np.int64)
    return hi, lo


def join_angle_delta(hi: np.ndarray, lo: np.ndarray) -> np.ndarray:
    hi = np.asarray(hi, np.float32)
    e = ((hi.view(np.uint32) >> 23) & 0xFF).astype(np.int64)
    return hi.astype(np.float64) + np.asarray(lo, np.float64) * np.ldexp(1.0, (e - 158).astype(np.int32))


def fresh_state_soa(x, y, vx, vy, angle, ang_vel, dry_mass, fuel, steps=0, ep_return=0.0,
                    fuel_unit: float = DEFAULT_FUEL_UNIT, dt: float = 0.1,
                    pos_unit: float = DEFAULT_POS_UNIT) -> np.ndarray:
    """A just-reset rocket (``Rocket.reset``, base.py:186): the kernel applies the Verlet
    bootstrap (base.py:65-130) itself on the first step.  All arguments are arrays [N]
    (or scalars); returns the public words as int32 [10, N].  ``dt``, ``fuel_unit`` and ``pos_unit``
    must be the environment's (``env.dt``, ``env.fuel_unit``, ``env.pos_unit``): the delta slots of a
    FRESH rocket hold ``v * dt``."""
    cols = np.broadcast_arrays(*[np.asarray(v, np.float64) for v in
                                 (x, y, angle, vx, vy, ang_vel, dry_mass, fuel)])
    cx, cy, cang, cvx, cvy, com, cdry, cfuel = [c.reshape(-1) for c in cols]
    n = cx.size
    soa = np.zeros((10, n), np.int32)
    soa[W_X], soa[W_Y] = _pos_units(cx, pos_unit), _pos_units(cy, pos_unit)
    soa[W_ANGLE] = _angle_units(cang)
    dt = float(dt)
    per_m = DELTA_UNITS_PER_POS_UNIT / pos_unit
    soa[W_SX] = np.rint(cvx * dt * per_m).astype(np.int32)      # v dt while FRESH: the first step
    soa[W_SY] = np.rint(cvy * dt * per_m).astype(np.int32)      # subtracts the bootstrap term a0 dt^2 / 2
    om_hi, om_lo = split_angle_delta(com * dt)                 # (omega dt with its 8 extra bits)
    soa[W_SANGLE] = _bits(om_hi)
    soa[W_DRY] = _bits(cdry)
    soa[W_FUEL] = _fuel_units(cfuel, fuel_unit)
    soa[W_META] = _meta_bits(np.broadcast_to(np.asarray(steps), (n,)), 0, True, om_lo)
    soa[W_EPRET] = _bits(np.broadcast_to(np.asarray(ep_return, np.float32), (n,)))
    return soa


def state_soa_from_reference(state: np.ndarray, prev: np.ndarray, steps, dt: float,
                             ep_return=0.0, fuel_unit: float = DEFAULT_FUEL_UNIT,
                             pos_unit: float = DEFAULT_POS_UNIT) -> np.ndarray:
    """Mid-episode injection from the reference's dicts.

    ``state`` [N,10] float64 in the reference's word order (x y vx vy ax ay angle
    angularVelocity mass fuelMass, i.e. ``info["raw_state"]``), ``prev`` [N,3] float64 =
    ``previous_state`` x, y, angle.  The carried angle delta is the PRE-wrap one
    (``angularVelocity * dt``, engine.py:220-222) and the wrap count is recovered from the
    normalised angles the reference stores."""
    state = np.asarray(state, np.float64)
    prev = np.asarray(prev, np.float64)
    n = state.shape[0]
    soa = np.zeros((10, n), np.int32)
    soa[W_X], soa[W_Y] = _pos_units(state[:, 0], pos_unit), _pos_units(state[:, 1], pos_unit)
    soa[W_ANGLE] = _angle_units(state[:, 6])
    per_m = DELTA_UNITS_PER_POS_UNIT / pos_unit
    soa[W_SX] = np.rint((state[:, 0] - prev[:, 0]) * per_m).astype(np.int32)
    soa[W_SY] = np.rint((state[:, 1] - prev[:, 1]) * per_m).astype(np.int32)
    pre_wrap = state[:, 7] * dt
    hi, lo = split_angle_delta(pre_wrap)
    soa[W_SANGLE] = _bits(hi)
    wraps = np.rint((pre_wrap - (state[:, 6] - prev[:, 2])) / 360.0).astype(np.int64)
    soa[W_DRY] = _bits(state[:, 8])
    soa[W_FUEL] = _fuel_units(state[:, 9], fuel_unit)
    soa[W_META] = _meta_bits(np.broadcast_to(np.asarray(steps), (n,)), wraps, False, lo)
    soa[W_EPRET] = _bits(np.broadcast_to(np.asarray(ep_return, np.float32), (n,)))
    return soa


def reference_view(soa: np.ndarray, dt: float, fuel_unit: float = DEFAULT_FUEL_UNIT,
                   pos_unit: float = DEFAULT_POS_UNIT) -> dict:
    """Public format -> the words of the reference's ``Rocket.state`` (float64 arrays) plus
    ``steps`` / ``fresh`` / ``wraps`` / ``ep_return``.  ``ax`` / ``ay`` are not persistent state
    (they are outputs of a step) and are not included.

    Velocities and the angle are formed in float32 exactly as the kernel forms the values it
    feeds its reward / termination tests (``float(units) * unit``, ``delta * unit * (1/dt)``), so
    that the reference's decision functions evaluated on this view agree with the kernel's flags
    bit for bit."""
    soa = as_words(soa)
    meta = soa[W_META].view(np.uint32)
    fresh = (meta & META_FRESH) != 0
    wraps = ((meta >> META_WRAP_SHIFT) & 7).astype(np.int64)
    wraps = np.where(wraps >= 4, wraps - 8, wraps)
    inv_dt = np.float32(1.0 / dt)
    unit = np.float32(pos_unit / DELTA_UNITS_PER_POS_UNIT)
    punit = np.float32(pos_unit)
    f = lambda w: soa[w].view(np.float32)                                   # noqa: E731

    def vel(w):                                   # (also while FRESH: the slots then hold v dt)
        return ((soa[w].astype(np.float32) * unit) * inv_dt).astype(np.float64)

    ang = soa[W_ANGLE].astype(np.float32) * np.float32(1.0 / ANGLE_UNITS_PER_DEG)
    om = (f(W_SANGLE) * inv_dt).astype(np.float64)
    return {"x": (soa[W_X].astype(np.float32) * punit).astype(np.float64),
            "y": (soa[W_Y].astype(np.float32) * punit).astype(np.float64),
            "vx": vel(W_SX), "vy": vel(W_SY), "angle": ang.astype(np.float64),
            "ang_vel": om, "dry_mass": f(W_DRY).astype(np.float64),
            "fuel": soa[W_FUEL].astype(np.float64) * fuel_unit,
            "steps": (meta & META_STEP_MASK).astype(np.int64),
            "fresh": fresh, "wraps": wraps, "ep_return": f(W_EPRET).astype(np.float64)}
