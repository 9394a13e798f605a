#!/usr/bin/env bash
# run on the GPU box: render throughput of fixture scenes under several values of one environment variable.  usage: VAR v1 v2 ... -- scene ...
var=$1; shift; vals=(); while [ "$1" != "--" ]; do vals+=("$1"); shift; done; shift
for v in "${vals[@]}"; do
  echo "== $var=$v"
  env $var=$v python tools/render_bench.py "$@" --size 1024 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %-18s paths %3d  %8.1f ms  %8.1f Mrays/s  rays %d' % (d['scene'], d['num_paths'], d['ms'], d['Mrays_per_s'], d['rays']))"
done
