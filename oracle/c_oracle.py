"""ctypes binding of the C oracle (``oracle/rocket_oracle.c``).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .rocket_oracle import EnvConstants, STATE_WORDS

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OracleConstants(C.Structure):
    _fields_ = ([(n, C.c_double) for n in (
        "dt", "gravity", "air_density", "thrust_power", "cold_gas_thrust_power", "fuel_rate",
        "radius", "moment_arm", "reference_area", "drag_coefficient", "angular_damping")]
        + [("reset_lo", C.c_double * 10), ("reset_hi", C.c_double * 10),
           ("land_perfect", C.c_double * 3), ("land_good", C.c_double * 3),
           ("land_ok", C.c_double * 3)]
        + [(n, C.c_double) for n in (
            "max_horizontal_position", "max_altitude", "tip_over_angle", "r_landing_perfect",
            "r_landing_good", "r_landing_ok", "r_crash_ground", "r_out_of_bounds",
            "r_tipped_over", "throttle_descent_reward_scale", "cold_gas_reward_scale",
            "free_fall_penalty_scale", "angle_aware_throttle_scale", "correct_direction_bonus",
            "gamma")]
        + [("obs_low", C.c_double * 8), ("obs_high", C.c_double * 8),
           ("max_episode_steps", C.c_int32), ("_pad", C.c_int32)])


ROCKET_DTYPE = np.dtype([(n, np.float64) for n in (
    "x", "y", "vx", "vy", "ax", "ay", "angle", "ang_vel", "dry_mass", "fuel",
    "prev_x", "prev_y", "prev_angle")] + [("current_step", np.int32), ("_pad", np.int32)])


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "rocket_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        assert L.oracle_sizeof_constants() == C.sizeof(OracleConstants)
        L.oracle_sizeof_rocket.restype = C.c_int64
        assert L.oracle_sizeof_rocket() == ROCKET_DTYPE.itemsize
        L.oracle_rollout_random.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                                            C.c_uint64, C.c_void_p]
        L.oracle_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 6
        L.oracle_bootstrap.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.oracle_observe.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.oracle_reward_only.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 5
        _LIB = L
    return _LIB


def pack_constants(c: EnvConstants) -> OracleConstants:
    k = OracleConstants()
    for n in ("dt", "gravity", "air_density", "thrust_power", "cold_gas_thrust_power",
              "fuel_rate", "radius", "moment_arm", "reference_area", "drag_coefficient",
              "angular_damping", "max_horizontal_position", "max_altitude", "tip_over_angle",
              "r_landing_perfect", "r_landing_good", "r_landing_ok", "r_crash_ground",
              "r_out_of_bounds", "r_tipped_over", "throttle_descent_reward_scale",
              "cold_gas_reward_scale", "free_fall_penalty_scale", "angle_aware_throttle_scale",
              "correct_direction_bonus", "gamma"):
        setattr(k, n, float(getattr(c, n)))
    for i, (lo, hi) in enumerate(c.reset_ranges):
        k.reset_lo[i], k.reset_hi[i] = lo, hi
    for i in range(3):
        k.land_perfect[i], k.land_good[i], k.land_ok[i] = (c.land_perfect[i], c.land_good[i],
                                                          c.land_ok[i])
    for i in range(8):
        k.obs_low[i], k.obs_high[i] = c.obs_low[i], c.obs_high[i]
    k.max_episode_steps = int(c.max_episode_steps)
    return k


class BatchOracle:
    """n independent rockets stepped by the C oracle; mirrors OracleEnv.inject/step."""

    def __init__(self, n: int, constants: EnvConstants | None = None):
        self.c = constants or EnvConstants.from_yaml()
        self._k = pack_constants(self.c)
        self.n = int(n)
        self.rockets = np.zeros(self.n, ROCKET_DTYPE)

    def inject(self, states: np.ndarray, current_step=0, mask=None) -> None:
        """states [n,10] in STATE_WORDS order (x y vx vy ax ay angle ang_vel dry fuel);
        recomputes the consistent previous state (base.py:65)."""
        states = np.asarray(states, np.float64)
        idx = np.arange(self.n) if mask is None else np.flatnonzero(mask)
        if idx.size == 0:
            return
        sub = np.zeros(idx.size, ROCKET_DTYPE)
        for j, w in enumerate(STATE_WORDS):
            sub[w] = states[idx, j]
        sub["current_step"] = (current_step if np.isscalar(current_step)
                               else np.asarray(current_step)[idx])
        lib().oracle_bootstrap(C.byref(self._k), sub.ctypes.data, sub.size)
        self.rockets[idx] = sub

    def state(self) -> np.ndarray:
        return np.stack([self.rockets[w] for w in STATE_WORDS], axis=1)

    def observe(self) -> np.ndarray:
        obs = np.zeros((self.n, 8), np.float32)
        lib().oracle_observe(C.byref(self._k), self.rockets.ctypes.data, self.n, obs.ctypes.data)
        return obs

    def step(self, actions: np.ndarray):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.n, 2)
        obs = np.zeros((self.n, 8), np.float32)
        rew = np.zeros(self.n, np.float64)
        term = np.zeros(self.n, np.uint8)
        trunc = np.zeros(self.n, np.uint8)
        land = np.zeros(self.n, np.int8)
        lib().oracle_step(C.byref(self._k), self.rockets.ctypes.data, self.n, a.ctypes.data,
                          obs.ctypes.data, rew.ctypes.data, term.ctypes.data,
                          trunc.ctypes.data, land.ctypes.data)
        return obs, rew, term.astype(bool), trunc.astype(bool), land


def reward_on_states(before: dict, after: dict, actions: np.ndarray,
                     constants: EnvConstants | None = None):
    """The reference's reward / termination / landing class (reward.py:27-281,
    lander.py:237-241) evaluated in fp64 on caller-supplied states.  ``before`` needs
    y, vx, vy, angle, ang_vel; ``after`` needs x, y, vx, vy, angle, ang_vel, steps."""
    k = pack_constants(constants or EnvConstants.from_yaml())
    n = len(after["y"])
    b = np.zeros(n, ROCKET_DTYPE)
    a = np.zeros(n, ROCKET_DTYPE)
    for w in ("y", "vx", "vy", "angle", "ang_vel"):
        b[w] = before[w]
    for w in ("x", "y", "vx", "vy", "angle", "ang_vel"):
        a[w] = after[w]
    a["current_step"] = after["steps"]
    act = np.ascontiguousarray(actions, np.float32).reshape(n, 2)
    rew = np.zeros(n, np.float64)
    term = np.zeros(n, np.uint8)
    trunc = np.zeros(n, np.uint8)
    land = np.zeros(n, np.int8)
    lib().oracle_reward_only(C.byref(k), b.ctypes.data, a.ctypes.data, n, act.ctypes.data,
                             rew.ctypes.data, term.ctypes.data, trunc.ctypes.data, land.ctypes.data)
    return rew, term.astype(bool), trunc.astype(bool), land


def rollout_random(n_envs_per_thread: int, steps: int, n_threads: int, seed: int = 0,
                   constants: EnvConstants | None = None) -> dict:
    """Random-action rollout with auto-reset on ``n_threads`` host threads (CPU baseline)."""
    k = pack_constants(constants or EnvConstants.from_yaml())
    out = (C.c_double * 9)()
    rc = lib().oracle_rollout_random(C.byref(k), n_envs_per_thread, steps, n_threads, seed, out)
    if rc != 0:
        raise RuntimeError("oracle_rollout_random failed")
    return {"env_steps": out[0], "seconds": out[1], "episodes": out[2], "sum_reward": out[3],
            "landings": [out[4 + i] for i in range(5)],
            "steps_per_s": out[0] / max(out[1], 1e-12)}
