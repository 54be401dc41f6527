#!/bin/bash
# One `ncu --set full` capture of the fused MLP kernel at 1 Mi rockets: `tools/mlp_ncu.sh <name> [skip]`.
# tools/mlp_probe.py launches, per batch size, 9 warm-up kernels, then 10 x policy net, 10 x value net, 10 x both
# in one launch, 10 x policy net with sampling; 1 Mi rockets is the third size: launches 98 .. 146, so
# skip 120 (default) = value net, 110 = policy net, 130 = mlp_pair_kernel.
cd ${GRAFT_REPO_ROOT:-.}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_ --launch-skip ${2:-120} --launch-count 1 -f -o gpurun_out/${1:-mlp} python tools/mlp_probe.py > gpurun_out/mlp_ncu.log 2>&1; tail -1 gpurun_out/mlp_ncu.log
