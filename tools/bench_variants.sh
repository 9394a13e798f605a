#!/usr/bin/env bash
# run on the GPU box: bench every build/variants/*.so (tools/tune.sh) under both traversal kernels and print the three launch times
# usage: tools/bench_variants.sh [warp|pool|flat|both] [lib-name-filter]
which=${1:-both}; filt=${2:-}
shopt -s nullglob
for lib in build/variants/*$filt*.so; do
  for k in warp pool flat; do
    [ "$which" != both ] && [ "$which" != $k ] && continue
    case $lib in *pool*|*d_s*) [ $k = warp ] && continue;; esac
    MIRO_GPU_LIB=$PWD/$lib MIRO_GPU_TRACE_KERNEL=$k python bench.py --steps 10 --warmup 3 --no-cpu --legs c2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$lib $k'.ljust(48), 'Mrays/s %7.1f  ms %s  frac %.3f' % (d['value'], ['%.3f'%x['ms'] for x in r['all_launches']], r['frac']))"
  done
done
