TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2g_n8.json 2> gpurun_out/r2g_n8.err
$TR --nproc-per-node 8 --master-port 29512 tools/pcie_ceiling.py > gpurun_out/r2g_pcie8.json 2> gpurun_out/r2g_pcie8.err
$TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2g_n4.json 2> gpurun_out/r2g_n4.err
$TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2g_n2.json 2> gpurun_out/r2g_n2.err
python bench.py --steps 20 --warmup 5 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r2g_n1.json 2> gpurun_out/r2g_n1.err
python tools/pcie_ceiling.py > gpurun_out/r2g_pcie1.json 2>&1
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1; lscpu | grep -i -E 'numa|socket|model name|^CPU\(s\)' >> gpurun_out/r2g_topo.txt
for f in n8 n4 n2 n1; do python -c "
import json; d=json.load(open('gpurun_out/r2g_$f.json')); print('$f', '%.4g'%d['value'], d['ms_per_step'], '%.3f'%d['roofline']['frac'], d['roofline']['kernel_ms'], d['clocks'].get('sm_mhz'), (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('copy_ceiling'))"; done
cat gpurun_out/r2g_pcie8.json gpurun_out/r2g_pcie1.json; tail -2 gpurun_out/r2g_n8.err
