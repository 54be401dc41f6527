#!/bin/bash
# Same-box A/B of step-kernel builds: `tools/ab_variants.sh name1 name2 ...` runs tools/step_probe.py on
# variants/lib_<name>.so (built by tools/build_variants.py, or copies of the in-tree library), interleaved
# and twice over -- the sustained clock under the board's power cap differs by a few per cent from box to
# box and drifts within a run, so only interleaved runs on one box compare builds -- then counts the
# warp-instructions of one steady-state launch of each with ncu (K = 1 and K = 8, 16 Mi rockets).
set -e
VARIANTS=${@:-$(ls variants/lib_*.so | sed 's/variants\/lib_\(.*\)\.so/\1/')}
for rep in 1 2; do
for v in $VARIANTS; do
  RL_B200_LIBRARY=$PWD/variants/lib_$v.so python tools/step_probe.py --rockets 67108864 --substeps 1 --launches 80 --repeat 2
  RL_B200_LIBRARY=$PWD/variants/lib_$v.so python tools/step_probe.py --rockets 16777216 --substeps 8 --launches 40 --repeat 2
done; done
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
for v in $VARIANTS; do
  RL_B200_LIBRARY=$PWD/variants/lib_$v.so ncu --metrics $M -k regex:step_kernel --launch-skip 2001 --launch-count 1 python tools/step_probe.py --rockets 16777216 --substeps 1 --launches 4 --repeat 1 2>&1 | grep -E "inst_executed|issue_active|duration" | sed "s/^/$v K1 /"
  RL_B200_LIBRARY=$PWD/variants/lib_$v.so ncu --metrics $M -k regex:step_kernel --launch-skip 251 --launch-count 1 python tools/step_probe.py --rockets 16777216 --substeps 8 --launches 4 --repeat 1 2>&1 | grep -E "inst_executed|issue_active|duration" | sed "s/^/$v K8 /"
done
