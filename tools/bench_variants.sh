#!/bin/bash
# Run bench.py's device leg on every variants/lib_*.so (GPU box), for every tiles-per-CTA value in $CHUNKS.
# usage: CHUNKS="0 1 4" tools/bench_variants.sh [rockets] [extra bench args]
ROCKETS=${1:-16777216}; shift
CHUNKS=${CHUNKS:-0}
mkdir -p gpurun_out
for lib in variants/lib_*.so; do
  name=$(basename $lib .so)
  for c in $CHUNKS; do
  RL_B200_TILES_PER_CTA=$c RL_B200_LIBRARY=$PWD/$lib python bench.py --rockets $ROCKETS --steps 100 --warmup 10 --no-e2e --no-cpu-baseline "$@" 2>gpurun_out/var_$name.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$name chunk=$c', '%.4g' % d['value'], '%.4f ms' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'], 'episodes', d.get('episodes'))"
  done
done | tee -a gpurun_out/variants.txt
