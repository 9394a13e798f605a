#!/usr/bin/env bash
# Renderer tuning helper: variants of libmiro_gpu.so whose render.cu is compiled with other -D knobs, into build/variants/
# usage: tools/tune_render.sh name "-DMIRO_SHADE_MIN_BLOCKS=6" [name2 "flags2" ...]
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG=$ROOT/rendering-algorithms-raytracer_b200
mkdir -p "$ROOT/build/variants"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  (
    tmp=$(mktemp -d)
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ccbin /usr/bin/g++ --compiler-options -fPIC,-ffp-contract=off -Xptxas -v $flags \
        -c "$PKG/csrc/render.cu" -o "$tmp/render.o" 2>&1 | grep -A2 "k_shadeILb0" | grep -E "registers|spill" | head -2
    /usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o "$ROOT/build/variants/$name.so" "$tmp/render.o" \
        "$PKG/build/miro_gpu_api.o" "$PKG/build/build.o" "$PKG/build/multi.o" "$PKG/build/miro_bvh.o" "$PKG/build/miro_host.o" "$PKG/build/miro_script.o" "$PKG/build/miro_host_capi.o"
    rm -rf "$tmp"; echo "built build/variants/$name.so ($flags)"
  ) &
done
wait
