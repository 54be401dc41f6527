"""``RocketSimulation``: the reference's UI-side caller of the step, batched on the GPU.

Mirrors the synchronous core of ``SimulationController`` (``backend/simulation/controller.py``:
``reset()``, ``step(actions)``) over ``RocketControls`` (``backend/rocket/controls.py``) and
produces, per tick, the exact binary message ``RocketWebSocketHandler.send_binary_telemetry``
(``backend/handler/websocket_handler.py:120-145``) writes to the socket: one header byte + 64 bytes
per rocket (``backend/protocol.py``).  The asyncio pacing loop, logging and the Tornado handler
are out of scope.
"""
from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np
import torch

from . import _lib
from .config import Config, RlParams, load_params

RECORD_FIELDS = ("x", "y", "vx", "vy", "ax", "ay", "angle", "angularVelocity", "angularAcceleration",
                 "mass", "fuelMass", "reward", "throttle", "coldGas", "landing_code", "is_active")
LANDING_NAMES = (None, "safe", "good", "ok", "unsafe")


class RocketSimulation:
    def __init__(self, num_rockets: int, device: Any = 0, seed: int = 0, config: Config | str | None = None,
                 params: RlParams | None = None):
        self._h = None
        self._L = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.RocketLibraryError("CUDA is not available: this simulation has no CPU fallback")
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        self.num_rockets = int(num_rockets)
        self.params = params if params is not None else load_params(config)
        self.dt = float(self.params.time_step)
        handle = C.c_void_p()
        _lib.check(self._L.rl_create(C.byref(self.params), self.num_rockets, 0, int(seed) & (2**64 - 1),
                                     dev.index or 0, None, C.byref(handle)), "rl_create")
        self._h = handle
        self._message = np.zeros(1 + 64 * self.num_rockets, np.uint8)
        self._obs = np.zeros((self.num_rockets, 8), np.float32)

    def reset(self) -> np.ndarray:
        """Re-arm every rocket (``SimulationController.reset``, controller.py:111-131).  Returns the
        clipped reset observations [N, 8] (the reference returns the state dicts)."""
        _lib.check(self._L.rl_reset_host(self._h, self._obs.ctypes.data), "rl_reset_host")
        return self._obs.copy()

    def set_state(self, state_words: np.ndarray) -> None:
        """Public state words, int32 [10, N] (``layout.fresh_state_soa``)."""
        w = np.asarray(state_words)
        if w.dtype != np.int32:
            raise TypeError("state words must be int32 [10, N]")
        t = torch.from_numpy(np.ascontiguousarray(w)).cuda()
        _lib.check(self._L.rl_set_state(self._h, t.data_ptr(), None, None), "rl_set_state")
        torch.cuda.synchronize()

    def step(self, actions: np.ndarray) -> bytes:
        """One tick for every rocket; returns the binary telemetry message (bytes)."""
        a = np.ascontiguousarray(actions, np.float32)
        if a.shape != (self.num_rockets, 2):
            raise ValueError(f"actions must have shape ({self.num_rockets}, 2)")
        _lib.check(self._L.rl_sim_step_host(self._h, a.ctypes.data, self._message.ctypes.data), "rl_sim_step_host")
        return self._message.tobytes()

    @staticmethod
    def decode(message: bytes) -> np.ndarray:
        """[N, 16] float32 view of a message (the decoder of frontend/src/services/telemetry.ts)."""
        if message[0] != 1 or (len(message) - 1) % 64:
            raise ValueError("not a telemetry message")
        return np.frombuffer(message, dtype="<f4", offset=1).reshape(-1, 16)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            self._L.rl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # noqa: BLE001
            pass
