set -x
B="python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
$B > gpurun_out/r2h_plain64.json 2>gpurun_out/r2h_plain64.err && ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 2008 --launch-count 1 -f -o gpurun_out/step_kernel_r2_64m $B > gpurun_out/r2h_ncu64.log 2>&1
K="python bench.py --rockets 16777216 --substeps 8 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
$K > gpurun_out/r2h_plaink8.json 2>gpurun_out/r2h_plaink8.err && ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 256 --launch-count 1 -f -o gpurun_out/step_kernel_r2_k8 $K > gpurun_out/r2h_ncuk8.log 2>&1
L="python bench.py --steps 20 --warmup 5 --spinup 100 --no-e2e --no-cpu-baseline --no-extras"
$L > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $L > gpurun_out/r2h_launches.log 2>&1
python tools/sanitizer_probe.py > gpurun_out/r2h_probe.log 2>&1     # (compute-sanitizer itself is closed on this pool: profiles/compute_sanitizer_r2.txt)
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2h_ref.json 2>gpurun_out/r2h_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -2 gpurun_out/r2h_probe.log
