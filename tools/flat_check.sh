#!/usr/bin/env bash
# run on the GPU box: the trace suites under the flat kernel, then the three workloads under each kernel back to back
# usage: tools/flat_check.sh [tests] [kernels...]
mkdir -p gpurun_out
if [ "$1" = tests ]; then shift
  python -m pytest tests/test_trace_gpu.py tests/test_synthetic_gpu.py -m gpu -x -q -k "flat or gives_the_warp or full_size" 2>&1 | tail -6
fi
for k in ${*:-warp flat}; do
  MIRO_GPU_TRACE_KERNEL=$k python bench.py --steps 20 --warmup 3 --no-cpu --legs ${LEGS:-c2,big,c5} 2>gpurun_out/flat_$k.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
rows=[('c2', d.get('value'), d.get('roofline'))]+[(k, v['Mrays_per_s'], v['roofline']) for k, v in d.get('workloads', {}).items()]
for name, val, r in rows:
    if r: print('$k'.ljust(6), name.ljust(6), 'Mrays/s %7.1f  ms %s  frac %.3f' % (val, ['%.3f'%x['ms'] for x in r['all_launches']], r['frac']))
" || tail -5 gpurun_out/flat_$k.err
done
