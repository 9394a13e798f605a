#!/usr/bin/env bash
# run on the GPU box: parity tests of one tuning build, then the bench of every variant
if [ -n "$1" ]; then MIRO_GPU_LIB=$PWD/build/variants/$1.so python -m pytest tests/test_trace_gpu.py tests/test_synthetic_gpu.py -m gpu -x -q 2>&1 | tail -3; fi
bash tools/bench_variants.sh
