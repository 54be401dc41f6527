#!/usr/bin/env python
"""Headline metrics + stall attribution per SASS line of the first kernel in an ncu report.
Usage: ncu_stalls_by_sass.py <report.ncu-rep> [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, r = rows[0], rows[2]
for w in ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
          'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg.per_second',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum']:
    if w in h:
        print(f'  {w:80s} {r[h.index(w)][:90]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, x in enumerate(rows) if x and x[0] == 'Kernel Name'] + [len(rows)]
blk = rows[starts[0] + 1: starts[1]]
h, data = blk[0], blk[1:]
ci = {n: i for i, n in enumerate(h)}
val = lambda x, k: float(x[ci[k]] or 0)    # noqa: E731
tot = sum(val(x, '# Samples') for x in data)
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
agg = {s: sum(val(x, s) for x in data) for s in stalls}
t = sum(agg.values()) or 1
print(f'warp samples {tot:.0f} over {len(data)} SASS lines; stall reasons (all samples): ' +
      ', '.join(f'{s[6:]}:{100 * v / t:.1f}%' for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
print('hottest SASS lines: index, samples, share, executed, instruction, top two stall reasons')
for idx, x in sorted(enumerate(data), key=lambda ix: -val(ix[1], '# Samples'))[:top]:
    t2 = sorted(((val(x, s), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{idx:5d} {val(x, '# Samples'):7.0f} {100 * val(x, '# Samples') / tot:5.1f}% {val(x, 'Instructions Executed'):9.0f}  "
          f"{x[ci['Source']][:64]:64s} {t2[0][1]}:{t2[0][0]:.0f} {t2[1][1]}:{t2[1][0]:.0f}")
