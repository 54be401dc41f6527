#!/bin/bash
# Same-box A/B of MLP-kernel builds: `tools/mlp_ab.sh name1 name2 ...` runs the rollout tests on the in-tree
# library, then tools/mlp_probe.py and tools/mlp_error_probe.py on variants/lib_<name>.so
# (tools/build_variants.py), interleaved, twice.
cd ${GRAFT_REPO_ROOT:-.}
timeout 600 python -m pytest tests/test_rollout.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2; do for v in "$@"; do echo "== $v"; RL_B200_LIBRARY=$PWD/variants/lib_$v.so timeout 120 python tools/mlp_probe.py; done; done
for v in "$@"; do echo "== error $v"; RL_B200_LIBRARY=$PWD/variants/lib_$v.so timeout 120 python tools/mlp_error_probe.py | sed -n 1p; done
