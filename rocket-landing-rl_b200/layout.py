"""Host-side helpers for the public state format of ``rl_set_state`` / ``rl_get_state``
(float32 ``[10][N]``, words ``RL_W_*`` of include/rocket_landing_b200.h).

The kernel does not store the reference's ``Rocket.state`` dict verbatim: it carries the
Verlet DELTAS ``x(t) - x(t-dt)`` (from which ``v = delta / dt``, engine.py:193-194) instead of
``previous_state`` (base.py:171), and a FRESH flag instead of a bootstrapped previous state
(base.py:65) -- see DESIGN.md "fp32 formulation".  These helpers convert in both directions,
doing the subtractions in float64 so that nothing is lost before the float32 rounding.
"""
from __future__ import annotations

import numpy as np

STATE_WORDS = ("x", "y", "angle", "sx", "sy", "sangle", "dry_mass", "fuel", "meta", "ep_return")
W_X, W_Y, W_ANGLE, W_SX, W_SY, W_SANGLE, W_DRY, W_FUEL, W_META, W_EPRET = range(10)
META_STEP_MASK = 0x00FFFFFF
META_WRAP_SHIFT = 24
META_FRESH = 0x80000000


def _meta_bits(steps, wraps, fresh) -> np.ndarray:
    steps = np.asarray(steps, np.int64) & META_STEP_MASK
    wraps = (np.asarray(wraps, np.int64) & 7) << META_WRAP_SHIFT
    bits = (steps | wraps | (META_FRESH if fresh else 0)).astype(np.uint32)
    return bits.view(np.float32)


def fresh_state_soa(x, y, vx, vy, angle, ang_vel, dry_mass, fuel, steps=0, ep_return=0.0) -> np.ndarray:
    """A just-reset rocket (``Rocket.reset``, base.py:186): the kernel applies the Verlet
    bootstrap (base.py:65-130) itself on the first step.  All arguments are arrays [N]
    (or scalars); returns float32 [10, N]."""
    cols = np.broadcast_arrays(*[np.asarray(v, np.float64) for v in
                                 (x, y, angle, vx, vy, ang_vel, dry_mass, fuel)])
    n = cols[0].size
    soa = np.zeros((10, n), np.float32)
    for w, col in zip((W_X, W_Y, W_ANGLE, W_SX, W_SY, W_SANGLE, W_DRY, W_FUEL), cols):
        soa[w] = col.reshape(n)
    soa[W_META] = _meta_bits(np.broadcast_to(np.asarray(steps), (n,)), 0, True)
    soa[W_EPRET] = np.broadcast_to(np.asarray(ep_return, np.float32), (n,))
    return soa


def state_soa_from_reference(state: np.ndarray, prev: np.ndarray, steps, dt: float,
                             ep_return=0.0) -> np.ndarray:
    """Mid-episode injection from the reference's dicts.

    ``state`` [N,10] float64 in the reference's word order (x y vx vy ax ay angle
    angularVelocity mass fuelMass, i.e. ``info["raw_state"]``), ``prev`` [N,3] float64 =
    ``previous_state`` x, y, angle.  The carried angle delta is the PRE-wrap one
    (``angularVelocity * dt``, engine.py:220-222) and the wrap count is recovered from the
    normalised angles the reference stores."""
    state = np.asarray(state, np.float64)
    prev = np.asarray(prev, np.float64)
    n = state.shape[0]
    soa = np.zeros((10, n), np.float32)
    soa[W_X], soa[W_Y], soa[W_ANGLE] = state[:, 0], state[:, 1], state[:, 6]
    soa[W_SX] = state[:, 0] - prev[:, 0]
    soa[W_SY] = state[:, 1] - prev[:, 1]
    pre_wrap = state[:, 7] * dt
    soa[W_SANGLE] = pre_wrap
    wraps = np.rint((pre_wrap - (state[:, 6] - prev[:, 2])) / 360.0).astype(np.int64)
    soa[W_DRY], soa[W_FUEL] = state[:, 8], state[:, 9]
    soa[W_META] = _meta_bits(np.broadcast_to(np.asarray(steps), (n,)), wraps, False)
    soa[W_EPRET] = np.broadcast_to(np.asarray(ep_return, np.float32), (n,))
    return soa


def reference_view(soa: np.ndarray, dt: float) -> dict:
    """Public format -> the words of the reference's ``Rocket.state`` (float64 arrays) plus
    ``steps`` / ``fresh`` / ``wraps`` / ``ep_return``.  ``ax`` / ``ay`` are not persistent state
    (they are outputs of a step) and are not included."""
    soa = np.asarray(soa, np.float32)
    meta = soa[W_META].view(np.uint32)
    fresh = (meta & META_FRESH) != 0
    wraps = ((meta >> META_WRAP_SHIFT) & 7).astype(np.int64)
    wraps = np.where(wraps >= 4, wraps - 8, wraps)
    inv_dt = np.float32(1.0 / dt)
    vel = lambda w: np.where(fresh, soa[w], soa[w] * inv_dt).astype(np.float64)   # noqa: E731
    return {"x": soa[W_X].astype(np.float64), "y": soa[W_Y].astype(np.float64),
            "vx": vel(W_SX), "vy": vel(W_SY), "angle": soa[W_ANGLE].astype(np.float64),
            "ang_vel": vel(W_SANGLE), "dry_mass": soa[W_DRY].astype(np.float64),
            "fuel": soa[W_FUEL].astype(np.float64), "steps": (meta & META_STEP_MASK).astype(np.int64),
            "fresh": fresh, "wraps": wraps, "ep_return": soa[W_EPRET].astype(np.float64)}
