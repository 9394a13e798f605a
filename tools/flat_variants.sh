#!/usr/bin/env bash
# run on the GPU box: the three workloads under the flat kernel for every build/variants/*.so (tools/tune.sh)
for lib in build/variants/*${1:-}*.so; do echo "== $lib"; MIRO_GPU_LIB=$PWD/$lib bash tools/flat_check.sh flat; done
