#!/usr/bin/env bash
# run on a multi-GPU box: tools/render_scale.py at N = 1, 2, 4 (, 8)
for n in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) tools/render_scale.py 2>/dev/null | grep '^{' | tee -a gpurun_out/render_scale.jsonl
done
