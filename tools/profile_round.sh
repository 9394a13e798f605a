#!/usr/bin/env bash
# run on the GPU box: tests, the bench line, then (each only after the plain command exited 0) the ncu launch list of the same
# command and one --set full capture of the three traversal launches of a step.  usage: tools/profile_round.sh TAG
tag=${1:-r1_v7}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { echo bench failed; tail -5 gpurun_out/bench_$tag.err; exit 1; }
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); r=d['roofline']
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', r['frac'], [a['ms'] for a in r['all_launches']], d['cpu_baseline']['value'], d['clocks'])"
python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 21 -c 3 -f -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep
