#!/usr/bin/env python
"""Join an ncu per-SASS-instruction table with nvdisasm line info and aggregate executed
warp-instructions per CUDA source line.  Usage:
    ncu_by_source_line.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> [rockets]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, lib, sub = sys.argv[1:4]
rockets = int(sys.argv[4]) if len(sys.argv) > 4 else 16777216
N = rockets / 32
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
sass = []
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith('.cubin')):   # the one that holds the kernel
    out = subprocess.run(['nvdisasm', '-g', '-c', cubin], cwd=tmp, capture_output=True, text=True).stdout
    if sub in out:
        sass = out.splitlines()
        break
# walk the kernel's section: remember the current "//## File ..., line N" for every instruction
lines, cur, inside = [], None, False
for ln in sass:
    if ln.startswith('.text.') and ln.endswith(':'):
        inside = sub in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
        lines.append((cur, ln.split('*/', 1)[1].strip().rstrip(';').strip()))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, x in enumerate(rows) if x and x[0] == 'Kernel Name'] + [len(rows)]
blk = rows[starts[0] + 1: starts[1]]
h, data = blk[0], blk[1:]
ci = {n: i for i, n in enumerate(h)}
assert len(data) == len(lines), (len(data), len(lines))
agg = collections.Counter(); samp = collections.Counter()
for (loc, text), row in zip(lines, data):
    agg[loc] += float(row[ci['Instructions Executed']] or 0)
    samp[loc] += float(row[ci['# Samples']] or 0)
tot = sum(agg.values()); ts = sum(samp.values())
print(f'total {tot / N:.1f} warp-instructions per warp-step')
srcs = {}
for (f, l), c in agg.most_common(int(os.environ.get('TOP', '45'))):
    if f not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), 'csrc', f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][l - 1].strip()[:70] if l - 1 < len(srcs[f]) else ''
    print(f'{c / N:7.1f} inst  {100 * samp[(f, l)] / ts:5.1f}% samples  {f}:{l:<4d} {text}')
