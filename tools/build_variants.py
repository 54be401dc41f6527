"""Build experiment variants of the product library (same sources, different -D knobs) into
variants/lib_<name>.so; bench them with tools/bench_variants.sh on the GPU box.

    python tools/build_variants.py name1:"-DRL_LOAD_MODE=1" name2:"-DRL_OBS_TRANSPOSE=1 -DRL_STAGES=3"
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "rocket-landing-rl_b200", "csrc")
OUT = os.path.join(ROOT, "variants")


def build(spec: str) -> str:
    name, _, flags = spec.partition(":")
    out = os.path.join(OUT, f"lib_{name}.so")
    cmd = ["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode",
           "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-Xptxas", "-v", *flags.split(),
           "-shared", "-o", out, "rl_capi.cu", "rl_vecnorm.cu", "rl_rollout.cu", "rl_policy.cu"]
    p = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if p.returncode != 0:
        return f"{name}: FAILED\n{p.stderr[-3000:]}"
    # registers of the headline instantiation step_kernel<true, true, DefaultParams>
    lines = p.stderr.splitlines()
    regs = "?"
    for k, line in enumerate(lines):
        if "step_kernel" in line and "Lb1ELb1ENS_13DefaultParams" in line and "Compiling" in line:
            for nxt in lines[k + 1:k + 6]:
                if "Used" in nxt:
                    regs = nxt.split("Used")[1].strip()
                    break
                if "spill" in nxt:
                    regs = nxt.strip() + " | "
            break
    return f"{name}: ok [{flags}] step_kernel<1,1,Default>: {regs}"


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    with ThreadPoolExecutor(max_workers=8) as ex:
        for msg in ex.map(build, sys.argv[1:]):
            print(msg)
