"""CPU tier: the bench.py contract that can be checked without a GPU -- the reference arm prints
one JSON line with the agreed keys, and the GPU arm refuses to run (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_one_json_line():
    p = _run("--impl", "reference", "--steps", "3", "--warmup", "1")
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["steps"] == 3 and d["warmup"] == 1
    cb = d["cpu_baseline"]
    # the reference itself (oracle/_ref or /root/reference) where it exists, else the C port
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["value"] > (2e3 * cb["cores"] if cb["kind"] == "reference" else 1e5)
    assert cb["c_port_steps_per_s"] > 1e5 and cb["python_oracle_steps_per_s_1core"] > 1e3
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_reference_arm_and_cuda_arm_state_the_same_config():
    """The driver compares the two arms' `config`: both come from workload_config()."""
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    a = argparse.Namespace(rockets=bench.TOTAL_ROCKETS, substeps=1, stats_interval=100)
    c = bench.workload_config(a, 1)
    assert "67108864 rockets" in c["workload"] and "l2" in c
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][0])["config"] == c


def test_byte_compiled_reference_travels_and_runs():
    """oracle/make_ref.py: the reference's env byte-compiled into ONE binary (no sources), unpacked
    and imported by ref_loader -- what times the reference on the GPU box."""
    if not os.path.isdir("/root/reference"):
        import pytest
        pytest.skip("needs the reference checkout (build container only)")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from oracle import make_ref; make_ref.build(quiet=True)\n"
            "from oracle.ref_loader import load_reference\n"
            "r = load_reference(prefer_compiled=True)\n"
            "assert 'rocket_ref_compiled_' in r.root, r.root\n"
            "e = r.RocketLandingEnv(); o, _ = e.reset(); out = e.step([0.5, 0.1])\n"
            "print(o.shape, len(out))\n") % ROOT
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert p.returncode == 0 and "(8,) 5" in p.stdout, p.stderr
    import zipfile
    names = zipfile.ZipFile(os.path.join(ROOT, "oracle", "_ref", "reference_backend.bin")).namelist()
    assert not [n for n in names if n.endswith(".py")] and "backend/envs/lander.pyc" in names


def test_reference_arm_other_ranks_exit_quietly():
    p = _run("--impl", "reference", "--steps", "2", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    p = _run("--steps", "1", "--warmup", "1")
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
