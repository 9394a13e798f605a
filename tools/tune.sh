#!/usr/bin/env bash
# Kernel tuning helper: builds variants of libmiro_gpu.so with different -D knobs into build/variants/ (they travel to the
# GPU box with the snapshot) — run them there with  MIRO_GPU_LIB=build/variants/<name>.so python bench.py --no-cpu ...
# usage: tools/tune.sh name "-DMIRO_TRACE_REFILL=4 ..." [name2 "flags2" ...]
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG=$ROOT/rendering-algorithms-raytracer_b200
mkdir -p "$ROOT/build/variants"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  (
    tmp=$(mktemp -d)
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ccbin /usr/bin/g++ --compiler-options -fPIC,-ffp-contract=off $flags \
        -c "$PKG/csrc/miro_gpu_api.cu" -o "$tmp/api.o"
    /usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o "$ROOT/build/variants/$name.so" "$tmp/api.o" \
        "$PKG/build/render.o" "$PKG/build/build.o" "$PKG/build/multi.o" "$PKG/build/miro_bvh.o" "$PKG/build/miro_host.o" "$PKG/build/miro_script.o" "$PKG/build/miro_host_capi.o"
    rm -rf "$tmp"; echo "built build/variants/$name.so ($flags)"
  ) &
done
wait
