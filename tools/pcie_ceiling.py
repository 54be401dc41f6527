#!/usr/bin/env python
"""Aggregate host<->device copy ceiling of the box: every rank copies between ITS GPU and pinned host
memory at the same time (what the e2e path of bench.py does on all ranks concurrently).

    python tools/pcie_ceiling.py                                  (one GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_ceiling.py

Prints one JSON line (rank 0): per-rank and aggregate GB/s for D2H alone, H2D alone and both directions
at once (the e2e step moves 38 B D2H + 8 B H2D per rocket), plus the NUMA node / CPU affinity of every
rank.  bench.py's `e2e.d2h_ceiling_gbs` is the same D2H measurement taken inside the bench run."""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = int(float(sys.argv[1]) * 2**30) if len(sys.argv) > 1 else 2**30
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_in = torch.empty(nbytes // 4, dtype=torch.uint8).pin_memory()
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_in = torch.empty(nbytes // 4, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=6):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        for s in (s1, s2):
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)       # the slowest rank bounds a synchronous step
        return float(t[0]) * 1e-3

    def d2h():
        with torch.cuda.stream(s1):
            h_out.copy_(d_out, non_blocking=True)

    def h2d():
        with torch.cuda.stream(s2):
            d_in.copy_(h_in, non_blocking=True)

    def both():
        d2h()
        h2d()

    t_d2h, t_h2d, t_both = timed(d2h), timed(h2d), timed(both)
    aff = sorted(os.sched_getaffinity(0))
    info = {"rank": rank, "cpus": f"{aff[0]}-{aff[-1]} ({len(aff)})"}
    try:
        prop = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
        info["gpu_numa_node"] = open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip()
    except Exception as e:   # noqa: BLE001
        info["gpu_numa_node"] = f"? ({type(e).__name__})"
    infos = [None] * world
    if world > 1:
        dist.all_gather_object(infos, info)
    else:
        infos = [info]
    if rank == 0:
        gb = nbytes / 1e9
        print(json.dumps({
            "ranks": world, "bytes_per_copy": nbytes,
            "d2h_gbs_per_rank": gb / t_d2h, "d2h_gbs_aggregate": world * gb / t_d2h,
            "h2d_gbs_per_rank": gb / 4 / t_h2d, "h2d_gbs_aggregate": world * gb / 4 / t_h2d,
            "both_d2h_gbs_aggregate": world * gb / t_both, "both_h2d_gbs_aggregate": world * gb / 4 / t_both,
            "e2e_ceiling_env_steps_per_s": world * nbytes / 38.0 / t_both,
            "placement": infos}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
