#!/usr/bin/env bash
# run on the GPU box: the round's final records — GPU tests, the full bench line + the reference arm, the ncu launch list of the bench
# command and ncu --set full pages of the traversal launches on c2 and big.   usage: tools/r2_final.sh TAG [tests bench launches ncu]
tag=${1:-r2_final}; shift
what=${*:-tests bench launches ncu}
mkdir -p gpurun_out
for w in $what; do
  case $w in
    tests) python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/gputests_$tag.log; tail -4 gpurun_out/gputests_$tag.log;;
    bench) python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$tag.err
           python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err; echo "ref rc=$?";;
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu --legs c2 > gpurun_out/ncu_launches_$tag.log 2>&1; wc -l gpurun_out/launches_$tag.csv;;
    ncu) bash tools/ncu_trace.sh ${tag}_c2; bash tools/r2_ncu.sh ${tag} big | tail -3;;
  esac
done
