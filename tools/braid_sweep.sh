#!/usr/bin/env bash
# run on the GPU box: instance braiding depth against the 40 401-instance field and the C5 render
for d in 0 1 2 3; do
  echo "== MIRO_BRAID_DEPTH=$d"
  MIRO_BRAID_DEPTH=$d python tools/instance_bench.py 201 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %-16s %7.1f Mrays/s  nodes %.1f tris %.1f insts %.1f' % (d['batch'], d['Mrays_per_s'], d['nodes_per_ray'], d['tris_per_ray'], d['instances_per_ray']))"
  MIRO_BRAID_DEPTH=$d python tools/render_bench.py c5_mb_instances --size 1024 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  render %-18s %8.1f ms  %8.1f Mrays/s' % (d['scene'], d['ms'], d['Mrays_per_s']))"
done
