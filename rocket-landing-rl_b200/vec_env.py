"""``RocketLandingSB3VecEnv``: the VecEnv surface the reference's rollout actually talks to.

``scripts/train.py:82-88`` builds ``make_vec_env(make_env, n_envs, SubprocVecEnv)`` and
wraps it in ``VecNormalize``; PPO's ``collect_rollouts`` then only ever calls
``reset() -> obs[N,8]`` and ``step(actions[N,2]) -> (obs[N,8], rewards[N], dones[N], infos)``
with numpy arrays, relies on auto-reset, ``infos[i]["terminal_observation"]``,
``infos[i]["TimeLimit.truncated"]`` and the ``Monitor`` record ``infos[i]["episode"] =
{"r", "l"}``.  This class offers exactly that surface with HOST (numpy) buffers on top
of ``rl_step_host`` -- it is the reference-facing plugin call whose throughput ``bench.py``
reports as ``e2e`` (host<->device copies inside the timed region).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from . import _lib
from .config import Config, RlParams, load_params
from .spaces import Box

_EMPTY: Dict[str, Any] = {}


def _pinned(shape, dtype) -> np.ndarray:
    t = torch.zeros(shape, dtype=dtype)
    try:
        t = t.pin_memory()
    except RuntimeError:      # no CUDA: plain pageable memory (construction fails later anyway)
        pass
    return t.numpy()          # the array keeps the (pinned) tensor storage alive




class RocketLandingSB3VecEnv:
    def __init__(self, num_envs: int, device: Any = 0, seed: int = 0,
                 config: Config | str | None = None, env_id_offset: int = 0, substeps: int = 1,
                 params: RlParams | None = None, build_infos: bool = True):
        self._h = None
        self._L = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.RocketLibraryError("CUDA is not available: this environment has no CPU fallback")
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        self.num_envs = int(num_envs)
        self.substeps = int(substeps)
        self.build_infos = bool(build_infos)
        self.params = params if params is not None else load_params(config)
        self.action_space = Box(low=np.array([0.0, -1.0], np.float32),
                                high=np.array([1.0, 1.0], np.float32), dtype=np.float32)
        self.observation_space = Box(low=np.array(list(self.params.obs_low), np.float32),
                                     high=np.array(list(self.params.obs_high), np.float32),
                                     dtype=np.float32)
        handle = C.c_void_p()
        _lib.check(self._L.rl_create(C.byref(self.params), self.num_envs, int(env_id_offset),
                                     int(seed) & (2**64 - 1), dev.index or 0, None, C.byref(handle)),
                   "rl_create")
        self._h = handle
        n = self.num_envs
        self._actions = _pinned((n, 2), torch.float32)
        self._obs = _pinned((n, 8), torch.float32)
        self._rew = _pinned((n,), torch.float32)
        self._term = _pinned((n,), torch.uint8)
        self._trunc = _pinned((n,), torch.uint8)
        if self.build_infos:      # per-episode records: only needed to build the SB3 info dicts
            self._tobs = _pinned((n, 8), torch.float32)
            self._epr = _pinned((n,), torch.float32)
            self._epl = _pinned((n,), torch.int32)
            self._land = _pinned((n,), torch.int8)
        self._pending = False

    # ---- SB3 VecEnv surface -------------------------------------------------------------
    def reset(self) -> np.ndarray:
        _lib.check(self._L.rl_reset_host(self._h, self._obs.ctypes.data), "rl_reset_host")
        return self._obs.copy()

    def step_async(self, actions: np.ndarray) -> None:
        a = np.asarray(actions, dtype=np.float32)
        if a.shape != (self.num_envs, 2):
            raise ValueError(f"actions must have shape ({self.num_envs}, 2), got {a.shape}")
        np.copyto(self._actions, a)
        self._pending = True

    def step_wait(self):
        if not self._pending:
            raise RuntimeError("step_wait() without step_async()")
        self._pending = False
        if self.build_infos:
            extra = (self._tobs.ctypes.data, self._epr.ctypes.data, self._epl.ctypes.data,
                     self._land.ctypes.data)
        else:
            extra = (None, None, None, None)
        _lib.check(self._L.rl_step_host(
            self._h, self._actions.ctypes.data, self._obs.ctypes.data, self._rew.ctypes.data,
            self._term.ctypes.data, self._trunc.ctypes.data, *extra, None, self.substeps, 1), "rl_step_host")
        dones = (self._term | self._trunc).astype(bool)
        infos: List[Dict[str, Any]] = self._infos(dones) if self.build_infos else []
        return self._obs.copy(), self._rew.copy(), dones, infos

    def step(self, actions: np.ndarray):
        self.step_async(actions)
        return self.step_wait()

    def step_raw(self, actions: np.ndarray):
        """Zero-copy variant: ``actions`` must be float32 [N,2]; returns views of the pinned
        output buffers (overwritten by the next call) and no per-rocket dicts."""
        _lib.check(self._L.rl_step_host(
            self._h, actions.ctypes.data, self._obs.ctypes.data, self._rew.ctypes.data,
            self._term.ctypes.data, self._trunc.ctypes.data, None, None, None, None, None,
            self.substeps, 1), "rl_step_host")
        return self._obs, self._rew, self._term, self._trunc

    def _infos(self, dones: np.ndarray) -> List[Dict[str, Any]]:
        infos: List[Dict[str, Any]] = [_EMPTY] * self.num_envs
        for i in np.flatnonzero(dones):
            infos[i] = {
                "terminal_observation": self._tobs[i].copy(),
                "TimeLimit.truncated": bool(self._trunc[i]) and not bool(self._term[i]),
                "episode": {"r": float(self._epr[i]), "l": int(self._epl[i])},
                "landing": ("none", "safe", "good", "ok", "unsafe")[int(self._land[i])],
            }
        return infos

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            self._L.rl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # noqa: BLE001
            pass

    # SB3 also probes these
    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * self.num_envs

    def get_attr(self, attr_name: str, indices=None):
        return [getattr(self, attr_name)] * self.num_envs

    def seed(self, seed: Optional[int] = None):
        return [None] * self.num_envs

    @property
    def launch_count(self) -> int:
        return int(self._L.rl_launch_count(self._h))
