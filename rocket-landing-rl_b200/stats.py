"""Episode statistics of a shard and their reduction across GPUs.

The step kernel accumulates a float64[16] vector per handle (``RL_S_*`` slots of
include/rocket_landing_b200.h).  Rockets never interact (reference: each ``Rocket`` owns its
dicts, base.py:26,49), so the ONLY cross-GPU traffic of the whole path is one
``all_reduce(SUM)`` of that 128-byte vector per reporting interval.
"""
from __future__ import annotations

from dataclasses import dataclass

SLOTS = ("episodes", "return_sum", "length_sum", "safe", "good", "ok", "unsafe", "time_limit",
         "out_of_bounds", "tip_steps", "nonfinite", "env_steps", "reward_sum")
STATS_WORDS = 16


@dataclass
class EpisodeStats:
    episodes: float = 0.0
    return_sum: float = 0.0
    length_sum: float = 0.0
    safe: float = 0.0
    good: float = 0.0
    ok: float = 0.0
    unsafe: float = 0.0
    time_limit: float = 0.0
    out_of_bounds: float = 0.0
    tip_steps: float = 0.0
    nonfinite: float = 0.0
    env_steps: float = 0.0
    reward_sum: float = 0.0

    @classmethod
    def from_vector(cls, vec) -> "EpisodeStats":
        vals = [float(v) for v in vec]
        return cls(**{name: vals[i] for i, name in enumerate(SLOTS)})

    @property
    def mean_return(self) -> float:          # SB3 Monitor's info["episode"]["r"], averaged
        return self.return_sum / self.episodes if self.episodes else float("nan")

    @property
    def mean_length(self) -> float:          # info["episode"]["l"], averaged
        return self.length_sum / self.episodes if self.episodes else float("nan")

    @property
    def landing_success_rate(self) -> float:  # safe + good + ok over finished episodes
        return (self.safe + self.good + self.ok) / self.episodes if self.episodes else float("nan")


def shard_range(total_envs: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous global rocket-id range ``[start, start + count)`` owned by ``rank``.  The
    first ``total % world`` ranks take one extra rocket.  ``start`` is the ``env_id_offset`` to
    pass to ``rl_create`` so that Philox resets are keyed by GLOBAL id and trajectories do
    not depend on the number of GPUs."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    if total_envs < world_size:
        raise ValueError(f"need at least one rocket per rank ({total_envs} < {world_size})")
    base, extra = divmod(total_envs, world_size)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def allreduce_stats(vec, group=None):
    """In-place SUM all-reduce of a statistics tensor (float64[16]) over ``group`` (NCCL on
    GPUs -- NVLink/NVSwitch, latency-bound; gloo in the CPU tests).  A no-op without an
    initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec
