cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/mlp_probe.py > gpurun_out/mlp_probe_r2c.json 2>&1; tail -1 gpurun_out/mlp_probe_r2c.json
timeout 300 python tools/rollout_bench.py --fused > gpurun_out/rollout_r2c.json 2>&1; tail -1 gpurun_out/rollout_r2c.json
bash tools/mlp_ncu.sh mlp_forward_r2c
