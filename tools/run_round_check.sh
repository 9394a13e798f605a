set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/gputests.log
cat gpurun_out/gputests.log
python bench.py > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; echo rc=$?
cat gpurun_out/bench_v7.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], [a['ms'] for a in d['roofline']['all_launches']], d['cpu_baseline'], d['render'])"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ref_v7.json 2>gpurun_out/ref_v7.err; tail -c 600 gpurun_out/ref_v7.json
