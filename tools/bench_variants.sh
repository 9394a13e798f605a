#!/usr/bin/env bash
# run on the GPU box: bench every build/variants/*.so (and the default lib) and print the three launch times
shopt -s nullglob
for lib in default build/variants/*.so; do
  if [ "$lib" = default ]; then unset MIRO_GPU_LIB; else export MIRO_GPU_LIB=$PWD/$lib; fi
  python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$lib'.ljust(28), 'Mrays/s %7.1f  ms %s  frac %.3f' % (d['value'], ['%.3f'%x['ms'] for x in r['all_launches']], r['frac']))"
done
