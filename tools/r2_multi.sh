#!/usr/bin/env bash
# run on a multi-GPU box (gpurun --gpus N): group tests, group strong scaling, the 2-rank NCCL test, bench.py under torchrun
# usage: tools/r2_multi.sh N [group|bench ...]
n=${1:-8}; shift
what=${*:-tests group bench}
for w in $what; do
  case $w in
  tests) python -m pytest tests/test_group_gpu.py tests/test_render_gpu.py -m gpu -q -k "group or nccl" 2>&1 | tail -5;;
  group) python tools/group_scale.py > gpurun_out/group_scale_r2.jsonl 2> gpurun_out/group_scale_r2.err; tail -3 gpurun_out/group_scale_r2.err; cut -c1-330 gpurun_out/group_scale_r2.jsonl;;
  bench)
    for k in 2 4 8; do
      [ $k -le $n ] || continue
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port $((29500 + k)) bench.py --gpus $k --steps 20 --warmup 3 > gpurun_out/bench_r2_${k}gpu.json 2> gpurun_out/bench_r2_${k}gpu.err
      echo "torchrun $k rc=$? lines=$(wc -l < gpurun_out/bench_r2_${k}gpu.json)"
      python -c "
import json; d=json.load(open('gpurun_out/bench_r2_${k}gpu.json')); print($k, round(d['value']), 'e2e', round(d['e2e']['value']), 'packed', round(d['e2e_packed']['value']), 'devcam', round(d['e2e_device_camera']['value']), d['host']); [print(r) for r in d.get('render_scaling', [])]"
    done;;
  esac
done
