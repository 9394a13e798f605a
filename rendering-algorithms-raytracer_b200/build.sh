#!/usr/bin/env bash
# Builds rendering-algorithms-raytracer_b200/libmiro_gpu.so IN-TREE: CUDA kernels for sm_100a + the C++ host layer,
# one shared library exporting the C ABIs of include/miro_gpu.h and include/miro_host.h.
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
HOSTCXX=${MIRO_CXX:-/usr/bin/g++}
OUT=$HERE/libmiro_gpu.so
OBJ=$HERE/build
mkdir -p "$OBJ"
CUFLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ccbin $HOSTCXX --compiler-options -fPIC,-ffp-contract=off ${MIRO_NVCC_EXTRA:-}"
pids=""
for f in miro_gpu_api render build multi; do
  $NVCC $CUFLAGS -c "$HERE/csrc/$f.cu" -o "$OBJ/$f.o" & pids="$pids $!"
done
for f in miro_bvh miro_host miro_script miro_host_capi; do
  $HOSTCXX -O2 -std=c++17 -fPIC -ffp-contract=off -Wall -Wno-unused-function -I/usr/local/cuda/include -c "$HERE/host/$f.cpp" -o "$OBJ/$f.o" & pids="$pids $!"
done
for p in $pids; do wait "$p"; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -ccbin $HOSTCXX -o "$OUT" "$OBJ"/miro_gpu_api.o "$OBJ"/render.o "$OBJ"/build.o "$OBJ"/multi.o "$OBJ"/miro_bvh.o "$OBJ"/miro_host.o "$OBJ"/miro_script.o "$OBJ"/miro_host_capi.o
# headless front end: scene script -> PPM, through the same library
$HOSTCXX -O2 -std=c++17 -I/usr/local/cuda/include "$HERE/host/miro_cli.cpp" -o "$HERE/miro_render" -L"$HERE" -lmiro_gpu -Wl,-rpath,'$ORIGIN' -L/usr/local/cuda/lib64 -lcudart
echo "built $OUT"
