"""``RocketLandingEnv``: ONE rocket behind exactly the surface of the reference's class of the
same name (``backend/envs/lander.py:18``) -- ``reset(*, seed=None) -> (obs float32[8], info)``,
``step(action float32[2]) -> (obs, reward float, terminated bool, truncated bool, info)``,
``action_space``, ``observation_space``, ``max_episode_steps``, ``current_step``, ``render()``,
``close()``; ``info`` carries the reference's keys (``raw_state`` with the ``Rocket.get_state()``
words, ``speed``, ``altitude``, ``steps``; lander.py:157-165).  No auto-reset: the caller resets
after ``terminated or truncated``, as gymnasium prescribes.

It exists so that code written against the reference's env -- ``make_env`` of
``scripts/train.py:48-50``, evaluation loops -- runs unchanged; each call is one launch of the same
CUDA kernel for a batch of one (about 20 us through the zero-copy host path), so anything that wants
throughput should use ``RocketLandingSB3VecEnv`` / ``RocketLandingVecEnv`` instead.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Any, Dict, Optional, Tuple

import numpy as np

from . import _lib
from .config import Config, RlParams, load_params
from .layout import reference_view
from .spaces import Box


class RocketLandingEnv:
    metadata: Dict[str, Any] = {"render_modes": []}

    def __init__(self, config: Config | str | None = None, seed: int = 0, device: int = 0,
                 params: RlParams | None = None):
        self._h = None
        self._L = _lib.load()
        self.params = params if params is not None else load_params(config)
        self.max_episode_steps = int(self.params.max_episode_steps)
        self.dt = float(self.params.time_step)
        self.current_step = 0
        self.action_space = Box(low=np.array([0.0, -1.0], np.float32), high=np.array([1.0, 1.0], np.float32),
                                dtype=np.float32)
        self.observation_space = Box(low=np.array(list(self.params.obs_low), np.float32),
                                     high=np.array(list(self.params.obs_high), np.float32), dtype=np.float32)
        handle = C.c_void_p()
        _lib.check(self._L.rl_create(C.byref(self.params), 1, 0, int(seed) & (2**64 - 1), int(device), None,
                                     C.byref(handle)), "rl_create")
        self._h = handle
        self.fuel_unit = float(self._L.rl_fuel_unit(self._h))
        self.pos_unit = float(self._L.rl_pos_unit(self._h))
        self._act = np.zeros((1, 2), np.float32)
        self._obs = np.zeros((1, 8), np.float32)
        self._rew = np.zeros((1,), np.float32)
        self._term = np.zeros((1,), np.uint8)
        self._trunc = np.zeros((1,), np.uint8)
        self._soa = np.zeros((10, 1), np.int32)
        self._raw_accel = np.zeros((1, 2), np.float32)
        self._accel = (0.0, 0.0)            # ax, ay of the last step (or the sampled ones after a reset)

    # ---- the reference's surface ------------------------------------------------------------
    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None) -> Tuple[np.ndarray, Dict[str, Any]]:
        """New random initial state (``seed`` is accepted and ignored, as in the reference: lander.py:170
        only seeds an unused generator; the stream here is fixed by the constructor's ``seed``)."""
        _lib.check(self._L.rl_reset_host(self._h, self._obs.ctypes.data), "rl_reset_host")
        self.current_step = 0
        obs = self._obs[0].copy()
        self._accel = (float(obs[4]), float(obs[5]))     # the sampled ax, ay: they only ever appear here (base.py:65-130)
        return obs, self._get_info()

    def step(self, action) -> Tuple[np.ndarray, float, bool, bool, Dict[str, Any]]:
        a = np.asarray(action, dtype=np.float32).reshape(2)
        self._act[0] = a
        _lib.check(self._L.rl_step_host(self._h, self._act.ctypes.data, self._obs.ctypes.data, self._rew.ctypes.data,
                                        self._term.ctypes.data, self._trunc.ctypes.data, None, None, None, None,
                                        self._raw_accel.ctypes.data, 1, 0), "rl_step_host")   # one substep, no auto-reset
        self.current_step += 1
        obs = self._obs[0].copy()
        # raw_state carries the UNCLIPPED accelerations of the step (lander.py:157-166; the observation clips to +-5)
        self._accel = (float(self._raw_accel[0, 0]), float(self._raw_accel[0, 1]))
        return obs, float(self._rew[0]), bool(self._term[0]), bool(self._trunc[0]), self._get_info()

    def _get_info(self) -> Dict[str, Any]:
        _lib.check(self._L.rl_get_state_host(self._h, self._soa.ctypes.data), "rl_get_state_host")
        v = reference_view(self._soa, self.dt, self.fuel_unit, self.pos_unit)
        vx, vy = float(v["vx"][0]), float(v["vy"][0])
        state = {"x": float(v["x"][0]), "y": float(v["y"][0]), "vx": vx, "vy": vy,
                 "ax": self._accel[0], "ay": self._accel[1], "angle": float(v["angle"][0]),
                 "angularVelocity": float(v["ang_vel"][0]), "mass": float(v["dry_mass"][0]),
                 "fuelMass": float(v["fuel"][0])}
        state["speed"] = math.sqrt(vx * vx + vy * vy)                   # Rocket.get_state, base.py:220-240
        state["relativeAngle"] = abs(state["angle"])
        state["totalMass"] = state["mass"] + state["fuelMass"]
        return {"raw_state": state, "speed": state["speed"], "altitude": state["y"], "steps": self.current_step}

    def render(self):
        """Placeholder, as in the reference (lander.py:257-260)."""
        return None

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            self._L.rl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # noqa: BLE001
            pass
