"""``RocketLandingVecEnv``: N rockets on one GPU behind the reference env's own surface.

Same names, argument meaning and spaces as the reference's ``RocketLandingEnv``
(``backend/envs/lander.py``): ``reset(*, seed=None) -> (obs, info)``,
``step(action) -> (obs, reward, terminated, truncated, info)``, ``action_space``,
``observation_space``, ``max_episode_steps``, ``render()``, ``close()`` -- with a leading
batch dimension and torch CUDA tensors as the zero-copy interface.  Auto-reset follows the
SB3 VecEnv contract the reference trains under (``scripts/train.py:82-88``).

All compute happens in ``librocket_landing_b200.so`` (hand-written sm_100a CUDA, loaded
through the C ABI).  There is no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .config import Config, RlParams, load_params
from .layout import reference_view
from .spaces import Box
from .stats import STATS_WORDS, EpisodeStats, allreduce_stats


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class RocketLandingVecEnv:
    """N independent rocket-landing environments stepped by one kernel launch.

    Parameters
    ----------
    num_envs       rockets on this GPU (this rank's shard)
    device         CUDA device (``"cuda:0"``, ``torch.device`` or ordinal)
    seed           Philox seed of the reset stream.  (The reference ignores ``seed`` --
                   lander.py:170 only seeds an unused generator, SURVEY.md Q11 -- here it
                   selects a reproducible counter-based stream instead.)
    config         ``Config`` / path of a ``config.yaml`` with the reference's key paths
    env_id_offset  global id of rocket 0 of this shard (multi-GPU sharding)
    substeps       K fused env steps per ``step`` call with the same action (rewards
                   summed, a rocket stops at its first done); 1 = the reference's step
    auto_reset     reset done rockets inside the step and return their reset observation
                   (SB3 VecEnv contract); the terminal observation is in
                   ``info["terminal_observation"]``
    raw_state      also report the step's UNCLIPPED accelerations as ``info["raw_accel"]`` [N,2]
                   (with ``get_reference_state()`` that is every word of the reference's
                   per-step ``info["raw_state"]``, lander.py:157-166; +8 B per rocket per step)
    """

    metadata: Dict[str, Any] = {"render_modes": []}

    def __init__(self, num_envs: int, device: Any = "cuda:0", seed: int = 0,
                 config: Config | str | None = None, env_id_offset: int = 0, substeps: int = 1,
                 auto_reset: bool = True, params: RlParams | None = None, raw_state: bool = False):
        self._h = None
        self._L = _lib.load()
        if isinstance(device, int):
            device = f"cuda:{device}"
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("RocketLandingVecEnv runs on CUDA devices only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise _lib.RocketLibraryError("CUDA is not available: this environment has no CPU fallback")
        self.num_envs = int(num_envs)
        if self.num_envs <= 0:
            raise ValueError("num_envs must be positive")
        self.substeps = int(substeps)
        self.auto_reset = bool(auto_reset)
        self.params = params if params is not None else load_params(config)
        self.dt = float(self.params.time_step)
        self.max_episode_steps = int(self.params.max_episode_steps)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)

        # spaces of ONE rocket, as the reference defines them (lander.py:65-69, :103-107)
        self.single_action_space = Box(low=np.array([0.0, -1.0], np.float32),
                                       high=np.array([1.0, 1.0], np.float32), dtype=np.float32)
        self.single_observation_space = Box(low=np.array(list(self.params.obs_low), np.float32),
                                            high=np.array(list(self.params.obs_high), np.float32),
                                            dtype=np.float32)
        self.action_space = self.single_action_space
        self.observation_space = self.single_observation_space

        n = self.num_envs
        nbytes = int(self._L.rl_state_bytes(n))
        # caller-owned state: one tensor = the whole persistent state (checkpoint = torch.save)
        self.state_buffer = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        handle = C.c_void_p()
        _lib.check(self._L.rl_create(C.byref(self.params), n, int(env_id_offset), int(seed) & (2**64 - 1),
                                     dev_index, self.state_buffer.data_ptr(), C.byref(handle)), "rl_create")
        self._h = handle
        self.fuel_unit = float(self._L.rl_fuel_unit(self._h))     # kg per unit of the fixed-point fuel word
        self.pos_unit = float(self._L.rl_pos_unit(self._h))       # m per unit of the fixed-point positions
        f32 = dict(dtype=torch.float32, device=self.device)
        self.obs = torch.zeros((n, 8), **f32)
        self.reward = torch.zeros((n,), **f32)
        self.terminated = torch.zeros((n,), dtype=torch.bool, device=self.device)
        self.truncated = torch.zeros((n,), dtype=torch.bool, device=self.device)
        self.terminal_obs = torch.zeros((n, 8), **f32)
        self.episode_return = torch.zeros((n,), **f32)
        self.episode_length = torch.zeros((n,), dtype=torch.int32, device=self.device)
        self.landing = torch.zeros((n,), dtype=torch.int8, device=self.device)
        self.raw_accel = torch.zeros((n, 2), **f32) if raw_state else None
        self._stats = torch.zeros((STATS_WORDS,), dtype=torch.float64, device=self.device)
        self._out_ptrs = None            # cached data_ptr()s of the output buffers (see step)

    # ------------------------------------------------------------------ gym surface
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None,
              mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Dict[str, Any]]:
        """Resample every rocket (or those with ``mask`` set) and return ``(obs [N,8], info)``
        (lander.py:168-186).  ``seed`` is accepted for API compatibility and ignored, exactly
        as in the reference; the Philox stream is fixed at construction."""
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _lib.check(self._L.rl_reset(self._h, _ptr(m), self.obs.data_ptr(), self._stream()), "rl_reset")
        return self.obs, {"steps": 0}

    def step(self, action: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor,
                                                  torch.Tensor, Dict[str, Any]]:
        """One env step for every rocket (lander.py:188-255).  ``action`` is float32 ``[N,2]``
        = (throttle, cold gas) on this device.  Returns the reference's 5-tuple, batched;
        the returned tensors are preallocated buffers overwritten by the next call."""
        if action.dtype != torch.float32 or action.device != self.device or not action.is_contiguous():
            action = action.to(device=self.device, dtype=torch.float32).contiguous()
        if action.shape != (self.num_envs, 2):
            raise ValueError(f"action must have shape ({self.num_envs}, 2), got {tuple(action.shape)}")
        outs = (self.obs, self.reward, self.terminated, self.truncated, self.terminal_obs,
                self.episode_return, self.episode_length, self.landing, self.raw_accel)
        if self._out_ptrs is None or self._out_ptrs[0] is not self.obs:      # buffers are persistent; a
            self._out_ptrs = (self.obs, tuple(_ptr(t) for t in outs))        # caller may swap self.obs
        rc = self._L.rl_step(self._h, action.data_ptr(), *self._out_ptrs[1], self.substeps,
                             int(self.auto_reset), self._stream())
        if rc != 0:
            _lib.check(rc, "rl_step")
        info = {"terminal_observation": self.terminal_obs, "episode_return": self.episode_return,
                "episode_length": self.episode_length, "landing": self.landing}
        if self.raw_accel is not None:
            info["raw_accel"] = self.raw_accel
        return self.obs, self.reward, self.terminated, self.truncated, info

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

    # ------------------------------------------------------------------ state injection / checkpoint
    def set_state(self, state_words, mask=None) -> None:
        """Inject persistent state in the public format (int32 words [10, N], see ``layout``)."""
        s = torch.as_tensor(state_words)
        if s.dtype != torch.int32:
            raise TypeError("state words must be int32 [10, N] (float words travel as their IEEE-754 bits; "
                            "see rocket_landing_rl_b200.layout)")
        s = s.to(self.device).contiguous()
        if s.shape != (10, self.num_envs):
            raise ValueError(f"state words must have shape (10, {self.num_envs})")
        m = None
        if mask is not None:
            m = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()
        _lib.check(self._L.rl_set_state(self._h, s.data_ptr(), _ptr(m), self._stream()), "rl_set_state")
        # s / m must outlive the enqueue: same-stream ordering keeps the memory valid until
        # the caching allocator reuses it on this stream, which is after the kernel.

    def get_state(self) -> torch.Tensor:
        """The persistent words of every rocket, int32 [10, N] (lossless: ``set_state`` of it restores
        the state bit for bit)."""
        out = torch.empty((10, self.num_envs), dtype=torch.int32, device=self.device)
        _lib.check(self._L.rl_get_state(self._h, out.data_ptr(), self._stream()), "rl_get_state")
        return out

    def raw_state(self) -> dict:
        """The reference's per-step ``info["raw_state"]`` (lander.py:157-166, ``Rocket.get_state``,
        base.py:220-240) for every rocket, as float64 numpy arrays keyed by the reference's names.
        ``ax`` / ``ay`` are the unclipped accelerations of the last step (needs ``raw_state=True``)."""
        if self.raw_accel is None:
            raise RuntimeError("construct the environment with raw_state=True")
        v = self.get_reference_state()
        acc = self.raw_accel.cpu().numpy().astype(np.float64)
        out = {"x": v["x"], "y": v["y"], "vx": v["vx"], "vy": v["vy"], "ax": acc[:, 0], "ay": acc[:, 1],
               "angle": v["angle"], "angularVelocity": v["ang_vel"], "mass": v["dry_mass"], "fuelMass": v["fuel"]}
        out["speed"] = np.sqrt(out["vx"] ** 2 + out["vy"] ** 2)
        out["relativeAngle"] = np.abs(out["angle"])
        out["totalMass"] = out["mass"] + out["fuelMass"]
        return out

    def get_reference_state(self) -> dict:
        """Persistent state as the reference's ``Rocket.state`` words (float64 numpy)."""
        return reference_view(self.get_state().cpu().numpy(), self.dt, self.fuel_unit, self.pos_unit)

    def state_dict(self) -> dict:
        return {"state": self.state_buffer.clone(), "epoch": int(self._L.rl_get_epoch(self._h)),
                "num_envs": self.num_envs}

    def load_state_dict(self, sd: dict) -> None:
        if sd["num_envs"] != self.num_envs:
            raise ValueError("checkpoint has a different num_envs")
        self.state_buffer.copy_(sd["state"])
        _lib.check(self._L.rl_set_epoch(self._h, int(sd["epoch"])), "rl_set_epoch")

    # ------------------------------------------------------------------ statistics
    def stats_tensor(self, reset: bool = False) -> torch.Tensor:
        """float64[16] statistics of THIS shard, on the device (enqueue only)."""
        _lib.check(self._L.rl_stats(self._h, self._stats.data_ptr(), int(reset), self._stream()), "rl_stats")
        return self._stats

    def stats_async(self, reset: bool = False, group=None) -> torch.Tensor:
        """Enqueue the statistics kernel and the 128-byte all-reduce over ``group``; returns the
        DEVICE tensor (float64[16]) without synchronising -- read it later with
        ``EpisodeStats.from_vector(t.cpu().tolist())``."""
        return allreduce_stats(self.stats_tensor(reset), group)

    def stats(self, reset: bool = False, group=None) -> EpisodeStats:
        """Episode statistics summed over all ranks of ``group`` (NCCL all-reduce of 128 B)."""
        vec = allreduce_stats(self.stats_tensor(reset), group)
        return EpisodeStats.from_vector(vec.cpu().tolist())

    @property
    def uses_folded_constants(self) -> bool:
        """True when the kernels with the reference's default constants folded in are used."""
        return bool(self._L.rl_uses_folded_constants(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.rl_launch_count(self._h))
