#!/usr/bin/env bash
# run on the GPU box: render throughput with overlapping waves (default) vs MIRO_GPU_RENDER_SERIAL=1, then the render tests
for mode in serial overlap serial overlap; do
  if [ $mode = serial ]; then export MIRO_GPU_RENDER_SERIAL=1; else unset MIRO_GPU_RENDER_SERIAL; fi
  echo "== $mode"
  python tools/render_bench.py c4_cornell_pt c3_dome_pt c5_mb_instances c9_texmaps --size 1024 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %-18s %4dx%-4d paths %3d  %8.1f ms  %8.1f Mrays/s  launches %d mean %.5f' % (d['scene'], d['size'][0], d['size'][1], d['num_paths'], d['ms'], d['Mrays_per_s'], d['kernel_launches'], d['mean']))"
done
unset MIRO_GPU_RENDER_SERIAL
python -m pytest tests/test_render_gpu.py -m gpu -x -q 2>&1 | tail -3
