"""Time the UNMODIFIED reference environment on this box's host cores.  MEASUREMENT INFRASTRUCTURE
ONLY (bench.py ``--impl reference`` and the ``cpu_baseline`` leg; never on the product path).

What BASELINE.md section 3 / SURVEY.md section 8d ask for: P = ``os.cpu_count()`` independent processes, each
one ``RocketLandingEnv`` (``backend/envs/lander.py:188``, imported through ``ref_loader`` -- on the GPU
box from the byte-compiled ``oracle/_ref``), uniform random actions, ``reset()`` on done, logging
disabled; warm-up, then a timed stretch; sum of steps / wall time of the slowest process.

The reference has no batch dimension: P processes stepping one rocket S times each do the work of one
"step" over a P x S-rocket sample of the benchmark's batch, which is how bench.py maps its
``--steps K --warmup W`` onto it.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import platform
import time


def _worker(idx: int, barrier, warm_steps: int, timed_steps: int, seconds: float, out) -> None:
    import numpy as np
    from oracle.ref_loader import load_reference
    ref = load_reference(allow_compiled=True)
    env = ref.RocketLandingEnv()
    np.random.seed(1000 + idx)
    rng = np.random.default_rng(idx)
    lo, hi = np.array([0.0, -1.0]), np.array([1.0, 1.0])
    env.reset()
    acts = rng.uniform(lo, hi, (4096, 2)).astype(np.float32)

    def run(n_steps, wall=None):
        t0 = time.perf_counter()
        k = 0
        while k < n_steps:
            _, _, term, trunc, _ = env.step(acts[k & 4095])
            if term or trunc:
                env.reset()
            k += 1
            if wall is not None and (k & 255) == 0 and time.perf_counter() - t0 >= wall:
                break
        return k, t0, time.perf_counter()

    run(warm_steps)
    barrier.wait()
    k, t0, t1 = run(timed_steps if seconds <= 0 else 1 << 62, wall=seconds if seconds > 0 else None)
    out.put((idx, k, t0, t1, ref.root))


def time_reference(processes: int | None = None, timed_steps: int = 0, warm_steps: int = 4000,
                   seconds: float = 0.0) -> dict:
    """Either ``timed_steps`` env steps per process, or a fixed wall of ``seconds`` per process."""
    procs = processes or os.cpu_count() or 1
    from oracle.ref_loader import reference_root
    if reference_root(allow_compiled=True) is None:            # (unpacks oracle/_ref once; the workers inherit it)
        raise RuntimeError("no reference to time: neither /root/reference nor oracle/_ref/reference_backend.bin")
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(procs)
    out = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(i, barrier, warm_steps, timed_steps, seconds, out)) for i in range(procs)]
    for p in ps:
        p.start()
    res = [out.get(timeout=600) for _ in ps]
    for p in ps:
        p.join()
    steps = sum(r[1] for r in res)
    wall = max(r[3] for r in res) - min(r[2] for r in res)     # CLOCK_MONOTONIC is system-wide
    cpu = platform.processor() or ""
    try:
        with open("/proc/cpuinfo") as fh:
            cpu = next((ln.split(":", 1)[1].strip() for ln in fh if ln.startswith("model name")), cpu)
    except OSError:
        pass
    return {"steps": steps, "seconds": wall, "steps_per_s": steps / wall, "processes": procs,
            "steps_per_s_per_core": steps / wall / procs, "cpu": cpu, "root": res[0][4]}


if __name__ == "__main__":
    import json
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    print(json.dumps(time_reference(seconds=float(sys.argv[1]) if len(sys.argv) > 1 else 5.0)))
