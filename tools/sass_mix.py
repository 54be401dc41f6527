"""Opcode histogram of one kernel of the built library (static SASS, `cuobjdump -sass`).

    python tools/sass_mix.py step_kernelILb1ELb1ELb1ENS_13DefaultParams [--top 40]

The argument is a substring of the MANGLED name (ILb1ELb1ELb1E = <auto-reset, single step, chunked>).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("RL_B200_LIBRARY") or os.path.join(ROOT, "rocket-landing-rl_b200", "librocket_landing_b200.so")


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else "step_kernelILb1ELb1ELb1ENS_13DefaultParams"
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = {}
    cur = None
    counts = collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            if cur not in names:
                names[cur] = subprocess.run(["cu++filt", cur], capture_output=True, text=True).stdout.strip()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            counts[cur][m.group(1).split(".")[0]] += 1
    for mangled, c in counts.items():
        if want in mangled:
            total = sum(c.values())
            print(f"{names[mangled]}\n  {total} static instructions")
            for op, k in c.most_common(top):
                print(f"  {op:12s} {k}")


if __name__ == "__main__":
    main()
