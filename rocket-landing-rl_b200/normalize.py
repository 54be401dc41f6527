"""``DeviceVecNormalize``: the reference's ``VecNormalize`` wrapper, on the device.

``scripts/train.py:86-88`` wraps the vectorised environment in
``VecNormalize(train_vec_env, norm_obs=True, norm_reward=True, clip_obs=10.0, gamma=GAMMA)``;
``backend/rl/agent.py:73-80,181`` re-loads its statistics to normalise observations at inference.
This class has the same constructor arguments and surface (``reset``, ``step``,
``normalize_obs``, ``get_original_obs``, ``get_original_reward``, ``training``) around a
``RocketLandingVecEnv``, with the statistics, the discounted returns and all arithmetic resident
on the GPU (``csrc/rl_vecnorm.cu`` through the C ABI).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Tuple

import numpy as np
import torch

from . import _lib
from .env import RocketLandingVecEnv


def read_sb3_vecnormalize(path: str) -> dict:
    """Statistics out of the ``vecnormalize.pkl`` that ``VecNormalize.save`` writes
    (``scripts/train.py:147``, re-loaded at ``backend/rl/agent.py:73-80``) without stable-baselines3
    or gymnasium installed: the pickle is read with stand-in classes for those two packages (numpy
    and builtins only otherwise), and only the running moments and the clip / epsilon / gamma
    settings are kept.  Returns what ``DeviceVecNormalize.load_stats`` takes, plus the settings."""
    import pickle

    class _Bag:
        def __init__(self, *a, **k):
            pass

        def __setstate__(self, state):
            self.__dict__.update(state if isinstance(state, dict) else {"state": state})

    # exact (module, name) pairs a VecNormalize pickle needs besides the stand-ins; anything else --
    # builtins.eval / exec / getattr / __import__ included -- is refused
    allowed = {("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.multiarray", "_reconstruct"),
               ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "_reconstruct"),
               ("numpy._core.multiarray", "scalar"), ("numpy.core.numeric", "_frombuffer"),
               ("numpy._core.numeric", "_frombuffer"), ("collections", "OrderedDict"),
               ("builtins", "tuple"), ("builtins", "list"), ("builtins", "dict"), ("builtins", "set"),
               ("builtins", "frozenset"), ("builtins", "slice"), ("builtins", "complex"),
               ("builtins", "bytearray"), ("builtins", "object")}

    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module.split(".")[0] in ("stable_baselines3", "gymnasium", "gym"):
                return _Bag
            if (module, name) in allowed:
                return super().find_class(module, name)
            raise pickle.UnpicklingError(f"refusing global {module}.{name} in {path}")

    with open(path, "rb") as fh:
        vn = _Unpickler(fh).load()
    return {"obs_mean": np.asarray(vn.obs_rms.mean, np.float64), "obs_var": np.asarray(vn.obs_rms.var, np.float64),
            "obs_count": float(vn.obs_rms.count), "ret_mean": float(vn.ret_rms.mean), "ret_var": float(vn.ret_rms.var),
            "ret_count": float(vn.ret_rms.count), "clip_obs": float(vn.clip_obs), "clip_reward": float(vn.clip_reward),
            "epsilon": float(vn.epsilon), "gamma": float(vn.gamma)}


class DeviceVecNormalize:
    def __init__(self, venv: RocketLandingVecEnv, training: bool = True, norm_obs: bool = True,
                 norm_reward: bool = True, clip_obs: float = 10.0, clip_reward: float = 10.0,
                 gamma: float = 0.99, epsilon: float = 1e-8):
        self._h = None
        self._L = _lib.load()
        self.venv = venv
        self.num_envs = venv.num_envs
        self.device = venv.device
        self.training, self.norm_obs, self.norm_reward = bool(training), bool(norm_obs), bool(norm_reward)
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = clip_obs, clip_reward, gamma, epsilon
        self.observation_space = venv.observation_space
        self.action_space = venv.action_space
        handle = C.c_void_p()
        rc = self._L.rl_vecnorm_create(self.num_envs, float(gamma), float(epsilon), float(clip_obs),
                                       float(clip_reward), self.device.index, C.byref(handle))
        if rc != 0:
            raise _lib.RocketLibraryError(
                f"rl_vecnorm_create failed ({rc}): {self._L.rl_vecnorm_last_error().decode()}")
        self._h = handle
        n = self.num_envs
        self.obs = torch.zeros((n, 8), dtype=torch.float32, device=self.device)
        self.reward = torch.zeros((n,), dtype=torch.float32, device=self.device)
        self.old_obs = venv.obs          # the wrapped env's own (unnormalised) output buffers
        self.old_reward = venv.reward

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise _lib.RocketLibraryError(f"{what} failed ({rc}): {self._L.rl_vecnorm_last_error().decode()}")

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def reset(self, **kwargs) -> torch.Tensor:
        raw, _ = self.venv.reset(**kwargs)
        self._check(self._L.rl_vecnorm_reset(self._h, raw.data_ptr(), self.obs.data_ptr(), int(self.training),
                                             int(self.norm_obs), self._stream()), "rl_vecnorm_reset")
        return self.obs

    def step(self, action: torch.Tensor, out_obs: torch.Tensor = None, out_reward: torch.Tensor = None
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, Dict[str, Any]]:
        """``VecNormalize.step``.  ``out_obs`` / ``out_reward`` (contiguous float32 ``[N, 8]`` / ``[N]``,
        e.g. slices of a rollout buffer) receive the normalised observation / reward instead of this
        wrapper's own buffers."""
        raw_obs, raw_rew, term, trunc, info = self.venv.step(action)
        obs = self.obs if out_obs is None else out_obs
        rew = self.reward if out_reward is None else out_reward
        assert obs.is_contiguous() and rew.is_contiguous() and obs.dtype == rew.dtype == torch.float32
        self._check(self._L.rl_vecnorm_step(
            self._h, raw_obs.data_ptr(), raw_rew.data_ptr(), term.data_ptr(), trunc.data_ptr(),
            info["terminal_observation"].data_ptr(), obs.data_ptr(), rew.data_ptr(),
            int(self.training), int(self.norm_obs), int(self.norm_reward), self._stream()), "rl_vecnorm_step")
        return obs, rew, term, trunc, info

    def normalize_obs(self, obs: torch.Tensor) -> torch.Tensor:
        """``VecNormalize.normalize_obs`` for an arbitrary ``[rows, 8]`` batch (inference path)."""
        obs = obs.to(device=self.device, dtype=torch.float32).contiguous()
        if not self.norm_obs:
            return obs
        out = torch.empty_like(obs)
        self._check(self._L.rl_vecnorm_normalize_obs(self._h, obs.data_ptr(), out.data_ptr(), obs.shape[0],
                                                     self._stream()), "rl_vecnorm_normalize_obs")
        return out

    def get_original_obs(self) -> torch.Tensor:
        return self.old_obs

    def get_original_reward(self) -> torch.Tensor:
        return self.old_reward

    # ---- statistics (what SB3 pickles as vecnormalize.pkl) -----------------------------------
    def stats(self) -> dict:
        buf = (C.c_double * 20)()
        self._check(self._L.rl_vecnorm_get_stats(self._h, buf), "rl_vecnorm_get_stats")
        v = np.array(buf[:])
        return {"obs_mean": v[0:8].copy(), "obs_var": v[8:16].copy(), "obs_count": float(v[16]),
                "ret_mean": float(v[17]), "ret_var": float(v[18]), "ret_count": float(v[19])}

    def load_stats(self, st: dict) -> None:
        v = np.concatenate([st["obs_mean"], st["obs_var"], [st["obs_count"], st["ret_mean"], st["ret_var"],
                                                              st["ret_count"]]]).astype(np.float64)
        buf = (C.c_double * 20)(*v.tolist())
        self._check(self._L.rl_vecnorm_set_stats(self._h, buf), "rl_vecnorm_set_stats")

    @property
    def launch_count(self) -> int:
        return int(self._L.rl_vecnorm_launch_count(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            self._L.rl_vecnorm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # noqa: BLE001
            pass
