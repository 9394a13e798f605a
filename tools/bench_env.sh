#!/usr/bin/env bash
# run on the GPU box: bench.py under several values of one environment variable.  usage: tools/bench_env.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$var=$v'.ljust(34), 'Mrays/s %7.1f e2e %7.1f  ms %s  frac %.3f nodes/ray %.2f tris/ray %.2f' % (d['value'], d['e2e']['value'], ['%.3f'%x['ms'] for x in r['all_launches']], r['frac'], r['nodes_per_ray'], r['tris_per_ray']))"
done
