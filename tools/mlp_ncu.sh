#!/bin/bash
# One `ncu --set full` capture of the value net's fused MLP kernel at 1 Mi rockets (launch 71 of tools/mlp_probe.py).
cd ${GRAFT_REPO_ROOT:-.}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_forward --launch-skip 70 --launch-count 1 -f -o gpurun_out/${1:-mlp} python tools/mlp_probe.py > gpurun_out/mlp_ncu.log 2>&1; tail -1 gpurun_out/mlp_ncu.log
