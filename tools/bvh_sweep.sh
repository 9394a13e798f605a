#!/usr/bin/env bash
# run on the GPU box: builder knobs against the C2 step
bash tools/bench_env.sh MIRO_BVH_BINS 16 32 64
bash tools/bench_env.sh MIRO_BVH_TRAVERSAL_COST 0.5 1 1.5 2
bash tools/bench_env.sh MIRO_BVH_MAX_LEAF 2 3 4 6
bash tools/bench_env.sh MIRO_BVH_ALPHA 0 1e-6 1e-4
