#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched rocket-landing environment step.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 (N = 1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json metric: "env-steps/sec at 1/2/4/8 B200 (64M rockets)"): 64 Mi rockets
in total, sharded by contiguous global id over the N ranks (strong scaling; no data-path
collective -- rockets never interact), one fresh random action per rocket per step, auto-reset
on, K_sub = 1 substep per launch.  A "step" is one launch of the step kernel over a rank's
whole shard.  The only collective is the NCCL all-reduce of the 16-word episode-statistics
vector every --stats-interval steps and at the end of the timed region.

Prints ONE JSON line (rank 0).  ``value`` = env-steps/s with everything resident in HBM;
``e2e`` = the same metric through the host-buffer VecEnv call (pinned numpy in/out, H2D and
D2H copies inside the timed region); ``roofline`` = algorithmic HBM bytes of the step kernel /
its CUDA-event duration against the measured copy bandwidth; ``cpu_baseline`` = the reference's
own environment (byte-compiled into ``oracle/_ref`` by ``oracle/make_ref.py``) on all host cores of
this box (N = 1 only), with the C port's and the Python oracle's rates beside it; ``extra`` = the
other BASELINE.json configurations measured after the headline (K = 8 fused substeps at 16 Mi
rockets, device-resident PPO rollout collection at 1 Mi rockets).  ``--impl reference`` times the
reference's environment only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOTAL_ROCKETS = 64 * 1024 * 1024
# algorithmic HBM bytes per env-step at one substep per launch (DESIGN.md "Roofline"):
#   state read 40 (planes A, B 16+16, fuel 4, dry 4) + state write 36 (dry is written on reset only)
#   + action read 8 + obs write 32 + reward write 4 + terminated/truncated 2 + landing class 1
BYTES_PER_ENV_STEP = 40 + 36 + 8 + 32 + 4 + 2 + 1
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


def _peaks() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:   # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:   # noqa: BLE001
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:   # noqa: BLE001
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


REF_STEPS_PER_BENCH_STEP = 1024     # env steps each reference process runs per bench "step"


def workload_config(args, world: int) -> dict:
    """The `config` object: identical for the CUDA arm and the reference arm."""
    per_launch = BYTES_PER_ENV_STEP * (args.rockets // world) / 1e9
    return {"workload": f"{args.rockets} rockets total sharded over {world} GPU(s), random actions, auto-reset in steady "
                        f"state, {args.substeps} substep(s) per launch, episode-statistics all-reduce every "
                        f"{args.stats_interval} steps",
            "l2": "inputs larger than L2: %.1f GB of state+action+output per launch per GPU vs 126 MB" % per_launch}


def cpu_port_rate(seconds_budget: float = 6.0, threads: int | None = None) -> dict:
    """The C port (oracle/rocket_oracle.c: scalar fp64 restatement of the reference's step, random
    actions, auto-reset) on all host cores, on a bounded sample of the workload."""
    from oracle import c_oracle
    threads = threads or os.cpu_count() or 1
    per_thread = 4096
    probe = c_oracle.rollout_random(per_thread, 50, threads, seed=1)
    steps = max(50, int(seconds_budget * probe["steps_per_s"] / (per_thread * threads)))
    out = c_oracle.rollout_random(per_thread, steps, threads, seed=2)
    return {"value": out["steps_per_s"], "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{per_thread * threads} rockets x {steps} steps, random actions, auto-reset, "
                      f"C fp64 port of the reference step on {threads} host threads ({out['seconds']:.1f} s)"}


def python_oracle_rate(steps: int = 5000) -> float:
    """Single-core rate of the Python ORACLE (oracle/rocket_oracle.py, a restatement -- NOT the
    reference; the reference itself is `cpu_baseline.value / cores`)."""
    import numpy as np
    from oracle import rocket_oracle as ro
    env = ro.OracleEnv()
    rng = np.random.default_rng(0)
    acts = rng.uniform([0, -1], [1, 1], (steps, 2)).astype(np.float32)
    t0 = time.perf_counter()
    for t in range(steps):
        _, _, term, trunc, _ = env.step(acts[t])
        if term or trunc:
            env.reset()
    return steps / (time.perf_counter() - t0)


def reference_available() -> bool:
    from oracle.ref_loader import reference_root
    return reference_root(allow_compiled=True) is not None


def cpu_baseline(seconds: float = 10.0) -> dict:
    """The reference's own RocketLandingEnv on every host core of this box (kind "reference"), or,
    where neither /root/reference nor oracle/_ref exists, the C port (kind "port")."""
    from oracle import c_oracle
    c_oracle.build()
    port = cpu_port_rate()
    extra = {"c_port_steps_per_s": port["value"], "c_port_threads": port["cores"],
             "python_oracle_steps_per_s_1core": python_oracle_rate()}
    if not reference_available():
        port.update(extra)
        return port
    from oracle import ref_timing
    r = ref_timing.time_reference(seconds=seconds)
    out = {"value": r["steps_per_s"], "unit": UNIT, "cores": r["processes"], "kind": "reference",
           "sample": f"{r['processes']} processes x one unmodified RocketLandingEnv (backend/envs/lander.py, from "
                     f"{os.path.relpath(r['root'], ROOT) if r['root'].startswith(ROOT) else r['root']}), random actions, reset on "
                     f"done, {r['steps']} env steps in {r['seconds']:.1f} s after a 4000-step warm-up each",
           "steps_per_s_per_core": r["steps_per_s_per_core"], "cpu": r["cpu"]}
    out.update(extra)
    return out


def run_reference(args) -> None:
    """`--impl reference`: the reference's own CPU implementation of the path on all host cores.
    One bench "step" = REF_STEPS_PER_BENCH_STEP env steps in each of P processes (a P x 1024-rocket
    sample of the benchmark's batch; the reference has no batch dimension)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle
    c_oracle.build()
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    S = REF_STEPS_PER_BENCH_STEP
    if reference_available():
        from oracle import ref_timing
        r = ref_timing.time_reference(timed_steps=args.steps * S, warm_steps=max(args.warmup, 1) * S)
        value, seconds, cores, kind = r["steps_per_s"], r["seconds"], r["processes"], "reference"
        sample = (f"{cores} processes x one unmodified RocketLandingEnv x {S} env steps per bench step "
                  f"({r['steps']} env steps in {seconds:.2f} s), random actions, reset on done; {r['cpu']}")
        dtype = "f64"
    else:
        threads = os.cpu_count() or 1
        if args.warmup > 0:
            c_oracle.rollout_random(S, args.warmup, threads, seed=1)
        out = c_oracle.rollout_random(S, args.steps, threads, seed=2)
        value, seconds, cores, kind = out["steps_per_s"], out["seconds"], threads, "port"
        sample = f"{S * threads} rockets x {args.steps} steps on {threads} host threads, C fp64 port (oracle/rocket_oracle.c)"
        dtype = "f64"
    port = cpu_port_rate(3.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * seconds / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "c_port_steps_per_s": port["value"], "c_port_threads": port["cores"],
                         "python_oracle_steps_per_s_1core": python_oracle_rate()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measure_extras(torch, dev) -> dict:
    """The other BASELINE.json configurations, after the headline (N = 1 only):
    configs[2] 16 Mi rockets with K = 8 fused substeps per launch (SM-issue-bound) and
    configs[4] device-resident PPO rollout collection at 1 Mi rockets (env -> VecNormalize -> fused
    tcgen05 policy + value nets -> [T,N] buffers -> GAE)."""
    from rocket_landing_rl_b200 import (ActorCriticMLP, DeviceRolloutBuffer, DeviceVecNormalize, FusedActorCritic,
                                        RocketLandingVecEnv, collect_rollout)
    out = {}
    n, K = 16 * 1024 * 1024, 8
    env = RocketLandingVecEnv(n, device=dev, seed=0, substeps=K)
    env.reset()
    gen = torch.Generator(device=dev).manual_seed(7)
    acts = []
    for _ in range(4):
        a = torch.rand((n, 2), device=dev, generator=gen)
        a[:, 1].mul_(2.0).sub_(1.0)
        acts.append(a)
    for t in range(260):                                   # 2080 env steps: episodes desynchronised
        env.step(acts[t % 4])
    env.stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    launches = 40
    e0.record()
    for t in range(launches):
        env.step(acts[t % 4])
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    st = env.stats()
    out["k8_16Mi"] = {"env_steps_per_s": st.env_steps / (ms * 1e-3), "ms_per_launch": ms / launches, "launches": launches,
                      "rockets": n, "substeps": K, "bound": "sm issue rate",
                      "issue_active_source": "profiles/ (ncu smsp__issue_active of the K = 8 capture, see profiles/README.md)"}
    env.close()
    del acts
    torch.cuda.empty_cache()

    n, T = 1 << 20, 32
    raw = RocketLandingVecEnv(n, device=dev, seed=0)
    venv = DeviceVecNormalize(raw, gamma=0.99)
    policy = FusedActorCritic(ActorCriticMLP().to(dev))
    buf = DeviceRolloutBuffer(T, n, dev)
    state = {"obs": venv.reset(), "starts": torch.ones(n, dtype=torch.bool, device=dev)}

    def one_pass():
        state["obs"], state["starts"] = collect_rollout(venv, policy, buf, state["obs"], state["starts"])

    for _ in range(4):
        one_pass()
    torch.cuda.synchronize(dev)
    passes = 3
    e0.record()
    for _ in range(passes):
        one_pass()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / passes
    out["rollout_1Mi"] = {"env_steps_per_s": n * T / (ms * 1e-3), "ms_per_collection_pass": ms, "rockets": n, "n_steps": T,
                          "pipeline": "rl_step -> rl_vecnorm_step -> fused tcgen05 policy + value nets (rl_policy.cu) -> "
                                      "[T,N] device buffers -> rl_gae"}
    raw.close()
    del buf, policy, venv, raw, state
    torch.cuda.empty_cache()

    # The denominator of roofline.frac is a BURST copy figure (MEASURED_PEAKS.json: best of 10 copies on an idle
    # board).  The headline kernel is timed after seconds of back-to-back launches, with the board on its power
    # cap; the same copy (torch b.copy_(a), 1 Gi bf16 elements, read + write bytes) is timed here both ways.
    a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev)
    b = torch.empty_like(a)
    nbytes = 2.0 * a.numel() * a.element_size()
    torch.cuda.synchronize(dev)
    time.sleep(0.5)                                          # let the board leave the cap
    burst = []
    for _ in range(10):
        e0.record()
        b.copy_(a)
        e1.record()
        torch.cuda.synchronize(dev)
        burst.append(e0.elapsed_time(e1))
    for _ in range(1500):                                    # ~1 s of back-to-back copies
        b.copy_(a)
    e0.record()
    for _ in range(200):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize(dev)
    out["copy_kernel"] = {"burst_gbs": nbytes / (min(burst) * 1e-3) / 1e9, "sustained_gbs": nbytes / (e0.elapsed_time(e1) / 200 * 1e-3) / 1e9,
                          "how": "torch b.copy_(a) over 1 Gi bf16 elements, read + write bytes: best of 10 on an idle board; "
                                 "200 copies after 1 500 back-to-back ones"}
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rockets", type=int, default=TOTAL_ROCKETS, help="total rockets over all ranks")
    ap.add_argument("--substeps", type=int, default=1)
    ap.add_argument("--stats-interval", type=int, default=100)
    ap.add_argument("--spinup", type=int, default=2000,
                    help="untimed env steps before the warm-up, so that episodes are desynchronised "
                         "(steady state: ~1/113 of the rockets reset per step, as in a long rollout)")
    ap.add_argument("--cooldown", type=float, default=0.0,
                    help="experiment only: idle this many seconds before the timed region (lets the board "
                         "leave its power cap)")
    ap.add_argument("--no-extras", action="store_true", help="skip the K = 8 and rollout-collection measurements")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from rocket_landing_rl_b200 import EpisodeStats, RocketLandingSB3VecEnv, RocketLandingVecEnv, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    start, count = shard_range(args.rockets, rank, world)
    warmup = max(args.warmup, 3)

    env = RocketLandingVecEnv(count, device=dev, seed=0, env_id_offset=start, substeps=args.substeps)
    env.reset()
    # a fresh random action per rocket per step: four pre-generated action sets, cycled
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_sets = 4
    actions = []
    for _ in range(n_sets):
        a = torch.rand((count, 2), device=dev, generator=gen)
        a[:, 1].mul_(2.0).sub_(1.0)
        actions.append(a)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # All rockets start their first episode together; a long rollout does not look like that.
    # Spin up (untimed) until episode phases are mixed, then do the W warm-up steps.
    for t in range((args.spinup + args.substeps - 1) // args.substeps):
        env.step(actions[t % n_sets])
    for t in range(warmup):
        env.step(actions[t % n_sets])
    env.stats(reset=True)
    launches0 = env.launch_count

    sampler = ClockSampler(local_rank)
    barrier()
    if args.cooldown > 0:
        time.sleep(args.cooldown)
    if rank == 0:
        sampler.start()
    ev0, ev1, evk = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    if world > 1:
        # The host threads leave the barrier above up to ~1 ms apart (8 processes on one VM), and the
        # closing all-reduce makes every rank wait for the last one to START: align the GPUs themselves
        # with one more (tiny, on-stream) all-reduce, so that every rank's clock starts within
        # microseconds of the others'.  It is part of the opening barrier, not of the timed region.
        align = torch.zeros(1, device=dev)
        dist.all_reduce(align)
    # Everything inside the timed region is ENQUEUED: K launches of the step kernel back to back on one
    # stream, the statistics kernel + 128-byte all-reduce (NCCL) every --stats-interval steps and once
    # at the end; the host reads the reduced statistics back after the closing event.
    ev0.record()
    for t in range(args.steps):
        env.step(actions[t % n_sets])
        if (t + 1) % args.stats_interval == 0 and t + 1 < args.steps:
            env.stats_async()
    evk.record()                                     # the K step-kernel launches end here
    final_vec = env.stats_async()                    # rl_stats + all-reduce, enqueued
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    final = EpisodeStats.from_vector(final_vec.cpu().tolist())
    elapsed_ms = ev0.elapsed_time(ev1)
    # the step kernel's own average launch duration (roofline): the launches run back to back on the
    # stream, so it is the span of the K launches / K -- no event pairs between launches (a pair around
    # a launch leaves the GPU idle for ~5 us); the in-loop statistics kernels (one per
    # --stats-interval steps, ~10 us) are inside this span.
    kernel_ms = ev0.elapsed_time(evk) / max(args.steps, 1)
    per_step = []
    launches = env.launch_count - launches0
    t_all = torch.tensor([elapsed_ms, kernel_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    elapsed_ms, kernel_ms = float(t_all[0]), float(t_all[1])
    # env steps actually executed (counted by the kernel; with substeps > 1 a rocket stops at its
    # first done, so this can be below rockets x steps x substeps)
    env_steps_total = float(final.env_steps)
    if args.substeps == 1:
        assert env_steps_total == float(args.rockets) * args.steps, (env_steps_total, args.rockets, args.steps)
    value = env_steps_total / (elapsed_ms * 1e-3)

    peak, peak_src = _peaks()
    bytes_per_launch = BYTES_PER_ENV_STEP * count / 1.0       # one launch = `count` rockets x substeps
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "rl::step_kernel<auto_reset, single_step=%s, DefaultParams>" % (args.substeps == 1), "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_env_step": BYTES_PER_ENV_STEP / args.substeps, "peak_source": peak_src}
    traffic_file = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fh:
                tj = json.load(fh)
            roofline["traffic"] = tj["dram_bytes_per_rocket"] * count
            roofline["traffic_source"] = tj.get("source")
        except Exception:   # noqa: BLE001
            pass
    env.close()
    del actions
    torch.cuda.empty_cache()

    # ---- e2e: the reference-facing VecEnv call with HOST buffers ----------------------------
    e2e = None
    if not args.no_e2e:
        import numpy as np
        henv = RocketLandingSB3VecEnv(count, device=local_rank, seed=0, env_id_offset=start,
                                      substeps=args.substeps, build_infos=False)
        henv.reset()
        hact = torch.rand((count, 2)).pin_memory()
        hact[:, 1].mul_(2.0).sub_(1.0)
        hnp = hact.numpy()
        henv.step_raw(hnp)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            obs, rew, term, trunc = henv.step_raw(hnp)
            _ = float(rew[0])                                # the result is on the host
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        launches += 0                                         # e2e launches are outside the timed region of `value`
        e2e = {"value": float(args.rockets) * args.e2e_steps * args.substeps / float(t_e[0]), "unit": UNIT,
               "h2d_bytes_per_step": 8 * count, "d2h_bytes_per_step": (32 + 4 + 1 + 1) * count,
               "steps": args.e2e_steps, "api": "RocketLandingSB3VecEnv.step_raw -> rl_step_host (pinned numpy in/out)"}
        henv.close()
        # The e2e path is bound by the host<->device copies, not by the kernel: measure the box's copy
        # ceiling for exactly these bytes -- every rank copies 38 B per rocket device->host and 8 B per
        # rocket host->device between ITS GPU and pinned memory, all ranks at the same time.
        d_out = torch.empty(38 * count, dtype=torch.uint8, device=dev)
        h_out = torch.empty(38 * count, dtype=torch.uint8).pin_memory()
        d_in = torch.empty(8 * count, dtype=torch.uint8, device=dev)
        s_out, s_in = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

        def copies():
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
            with torch.cuda.stream(s_in):
                d_in.copy_(hact.view(torch.uint8).reshape(-1), non_blocking=True)

        copies()
        barrier()
        # a CEILING: the best of five repetitions (one repetition = this rank's bytes of one step, both directions at
        # once); a mean over a few copies has come out below what rl_step_host then sustained on the same box
        best = float("inf")
        cur = torch.cuda.current_stream(dev)
        for _ in range(5):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            barrier()                                        # every repetition: all ranks copy at the same time
            c0.record()
            s_out.wait_stream(cur)
            s_in.wait_stream(cur)
            copies()
            cur.wait_stream(s_out)
            cur.wait_stream(s_in)
            c1.record()
            torch.cuda.synchronize(dev)
            t_rep = torch.tensor([c0.elapsed_time(c1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t_rep, op=dist.ReduceOp.MAX)    # the slowest rank of THIS repetition
            best = min(best, float(t_rep[0]))
        t_c = torch.tensor([best], device=dev, dtype=torch.float64)
        ceiling = float(args.rockets) * args.substeps / (float(t_c[0]) * 1e-3)
        e2e["copy_ceiling"] = {"env_steps_per_s": ceiling, "d2h_gbs_aggregate": 38.0 * args.rockets / (float(t_c[0]) * 1e-3) / 1e9,
                               "h2d_gbs_aggregate": 8.0 * args.rockets / (float(t_c[0]) * 1e-3) / 1e9,
                               "how": "all ranks at once: cudaMemcpyAsync of 38 B/rocket D2H + 8 B/rocket H2D, pinned memory, "
                                      "two streams, slowest rank of a repetition, best of five repetitions"}
        e2e["frac_of_copy_ceiling"] = e2e["value"] / ceiling
        del d_out, h_out, d_in

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": elapsed_ms / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "setup": f"{count} rockets on this rank; 4 pre-generated action sets cycled; {args.spinup} untimed spin-up "
                     f"steps desynchronise the episodes before the warm-up",
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "episodes": final.episodes, "mean_episode_length": final.mean_length,
        }
        if world == 1 and not args.no_extras:
            line["extra"] = measure_extras(torch, dev)
            # (informational, next to the contract's frac: the same kernel time against the copy rate this board sustains)
            roofline["frac_of_sustained_copy"] = roofline["achieved"] / line["extra"]["copy_kernel"]["sustained_gbs"]
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
