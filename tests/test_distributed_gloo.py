"""CPU tier, world_size 2 over gloo: the multi-GPU host logic -- each rank owns a contiguous
global rocket-id range, the only collective is the SUM all-reduce of the 16-word statistics
vector, and Philox resets keyed by GLOBAL id make the union of the shards identical to an
unsharded batch (checked on the host build of the device math)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rocket_landing_rl_b200 as pkg


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, total: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
        import parity
        start, count = pkg.shard_range(total, rank, world)
        st = parity.HostEmuStepper(count)
        st.philox_reset(gid0=start, seed=42, epoch=3)          # this rank's shard of the reset
        np.save(os.path.join(out_dir, f"shard{rank}.npy"), st.soa)
        # local statistics vector, then the path's single collective
        vec = torch.zeros(16, dtype=torch.float64)
        vec[0] = count                 # pretend every rocket finished one episode
        vec[1] = float(st.soa[1].astype(np.float64).sum())
        vec[11] = count * 10
        pkg.allreduce_stats(vec)
        stats = pkg.EpisodeStats.from_vector(vec.tolist())
        assert stats.episodes == total and stats.env_steps == total * 10
        torch.save(vec, os.path.join(out_dir, f"stats{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_stats_allreduce(tmp_path):
    total, world = 1001, 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    import parity
    whole = parity.HostEmuStepper(total)
    whole.philox_reset(gid0=0, seed=42, epoch=3)
    shards = [np.load(tmp_path / f"shard{r}.npy") for r in range(world)]
    np.testing.assert_array_equal(np.concatenate(shards, axis=1), whole.soa)
    v0, v1 = torch.load(tmp_path / "stats0.pt"), torch.load(tmp_path / "stats1.pt")
    assert torch.equal(v0, v1)
    assert v0[1].item() == np.float64(whole.soa[1].astype(np.float64).sum())


def test_allreduce_is_noop_without_process_group():
    vec = torch.arange(16, dtype=torch.float64)
    assert torch.equal(pkg.allreduce_stats(vec.clone()), vec)
