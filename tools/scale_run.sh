# 1 -> 8 GPU strong scaling of bench.py on one node, as the driver launches it (gpurun --gpus 8 -- bash tools/scale_run.sh)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --steps 20 --warmup 5 --no-e2e --no-extras --no-cpu-baseline 2> gpurun_out/scale_n1.err | grep '^{' > gpurun_out/scale_n1.json
$TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e 2> gpurun_out/scale_n2.err | grep '^{' > gpurun_out/scale_n2.json
$TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e 2> gpurun_out/scale_n4.err | grep '^{' > gpurun_out/scale_n4.json
$TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 2> gpurun_out/scale_n8.err | grep '^{' > gpurun_out/scale_n8.json
$TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --steps 200 --warmup 20 --no-e2e 2> gpurun_out/scale_n8_long.err | grep '^{' > gpurun_out/scale_n8_long.json
for f in n1 n2 n4 n8 n8_long; do python -c "
import json; d=json.load(open('gpurun_out/scale_$f.json')); e=d.get('e2e') or {}
print('$f', '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), 'e2e', e.get('value'), e.get('frac_of_copy_ceiling'))"; done
