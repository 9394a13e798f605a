#!/usr/bin/env bash
# run on the GPU box: one `ncu --set full` capture of the three traversal launches of a bench step (primary, incoherent, shadow)
# under the environment given after the tag; the report is turned into CSV pages (raw + SASS source) and removed.
# usage: tools/ncu_trace.sh TAG [VAR=value ...]
tag=$1; shift
env "$@" python bench.py --steps 2 --warmup 3 --no-cpu --legs c2 > gpurun_out/ncu_pre_$tag.json 2> gpurun_out/ncu_pre_$tag.err || { echo "bench failed"; tail -5 gpurun_out/ncu_pre_$tag.err; exit 1; }
env "$@" ncu --set full --clock-control none --import-source on -k regex:k_trace -s 20 -c 3 -f -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 3 --no-cpu --legs c2 > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_${tag}_source.csv 2>/dev/null
gzip -f gpurun_out/prof_${tag}_source.csv; rm -f gpurun_out/prof_$tag.ncu-rep
ls -la gpurun_out/prof_${tag}_raw.csv
