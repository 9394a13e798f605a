#!/usr/bin/env bash
# run on the GPU box (round 2): GPU tests, the parity table, the full bench line, then ncu captures.  usage: tools/r2_run.sh TAG [what...]
tag=${1:-r2}; shift
what=${*:-tests parity bench ncu}
for w in $what; do
  case $w in
    tests) python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/gputests_$tag.log; tail -8 gpurun_out/gputests_$tag.log;;
    parity) python tools/parity_report.py $tag > gpurun_out/parity_$tag.log 2>&1; tail -5 gpurun_out/parity_$tag.log;;
    bench) python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err
           python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err; echo "ref rc=$?";;
    ncu) bash tools/ncu_trace.sh ${tag}_warp MIRO_GPU_TRACE_KERNEL=warp; bash tools/ncu_trace.sh ${tag}_pool MIRO_GPU_TRACE_KERNEL=pool;;
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu --legs c2 > gpurun_out/ncu1.log 2>&1;;
  esac
done
