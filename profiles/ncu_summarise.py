#!/usr/bin/env python
"""Summarise an ncu report of the step kernel: headline metrics, opcode mix per warp-step,
stall reasons and the hottest SASS lines.  Usage: ncu_summarise.py <report.ncu-rep> [rockets]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
rockets = int(sys.argv[2]) if len(sys.argv) > 2 else 16777216
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.avg.per_second',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
r = rows[2]
print('kernel:', r[h.index('Kernel Name')][:90])
for w in WANT:
    if w in h:
        print(f'  {w:72s} {r[h.index(w)]:>18s} {rows[1][h.index(w)]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, x in enumerate(rows) if x and x[0] == 'Kernel Name'] + [len(rows)]
blk = rows[starts[0] + 1: starts[1]]
h, data = blk[0], blk[1:]
ci = {n: i for i, n in enumerate(h)}
N = rockets / 32
val = lambda x, k: float(x[ci[k]] or 0)
tot = sum(val(x, 'Instructions Executed') for x in data)
always = sum(val(x, 'Instructions Executed') for x in data if val(x, 'Instructions Executed') >= 0.999 * N)
print(f'warp-instructions per warp-step: {tot / N:.1f}  (always-executed lines: {always / N:.1f}, rest {tot / N - always / N:.1f}); SASS lines {len(data)}')
op = collections.Counter()
for x in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', x[ci['Source']])
    op[m.group(2).split('.')[0] if m else '?'] += val(x, 'Instructions Executed')
print('opcode mix:', ', '.join(f'{o}:{c / N:.1f}' for o, c in op.most_common(36)))
stalls = [n for n in h if n.startswith('stall_')]
agg = {s: sum(val(x, s) for x in data) for s in stalls}
t = sum(agg.values()) or 1
print('stalls:', ', '.join(f'{s[6:]}:{100 * v / t:.1f}%' for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
print('hottest SASS lines (samples, executed, text):')
for x in sorted(data, key=lambda x: -val(x, '# Samples'))[:14]:
    print('  ', x[ci['# Samples']], x[ci['Instructions Executed']], x[ci['Source']][:84])
