set -e
for rep in 1 2; do
for v in opt2 posfix2; do
  RL_B200_LIBRARY=$PWD/variants/lib_$v.so python tools/step_probe.py --rockets 67108864 --substeps 1 --launches 80 --repeat 2
  RL_B200_LIBRARY=$PWD/variants/lib_$v.so python tools/step_probe.py --rockets 16777216 --substeps 8 --launches 40 --repeat 2
done; done
for v in opt2 posfix2; do
  RL_B200_LIBRARY=$PWD/variants/lib_$v.so ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum -k regex:step_kernel --launch-skip 2001 --launch-count 1 python tools/step_probe.py --rockets 16777216 --substeps 1 --launches 4 --repeat 1 2>&1 | grep -E "inst_executed|issue_active|duration" | sed "s/^/$v K1 /"
  RL_B200_LIBRARY=$PWD/variants/lib_$v.so ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum -k regex:step_kernel --launch-skip 251 --launch-count 1 python tools/step_probe.py --rockets 16777216 --substeps 8 --launches 4 --repeat 1 2>&1 | grep -E "inst_executed|issue_active|duration" | sed "s/^/$v K8 /"
done
