#!/usr/bin/env bash
# run on the GPU box: ncu --set full of the three traversal launches of the `big` and `c5` workloads, of the shading / resolve
# kernels of one C4 frame, and the launch list of that frame.  Reports are turned into CSV pages on the box and removed (the
# merge back is limited to 64 MiB).  usage: tools/r2_ncu.sh TAG [big c5 render]
tag=${1:-r2}; shift
what=${*:-big c5 render}
pages() {   # rep-basename
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass > gpurun_out/$1_source.csv 2>/dev/null
  gzip -f gpurun_out/$1_source.csv
  rm -f gpurun_out/$1.ncu-rep
}
for w in $what; do
  case $w in
  big|c5)
    python bench.py --steps 2 --warmup 3 --no-cpu --legs $w > gpurun_out/pre_${tag}_$w.json 2> gpurun_out/pre_${tag}_$w.err || { echo "bench $w failed"; tail -3 gpurun_out/pre_${tag}_$w.err; continue; }
    ncu --set full --clock-control none --import-source on -k regex:k_trace -s 20 -c 3 -f -o gpurun_out/prof_${tag}_$w python bench.py --steps 2 --warmup 3 --no-cpu --legs $w > gpurun_out/ncu_${tag}_$w.log 2>&1
    pages prof_${tag}_$w;;
  render)
    python tools/render_bench.py c4_cornell_pt --size 512 --reps 1 > gpurun_out/pre_${tag}_render.json 2>&1 || { echo "render_bench failed"; continue; }
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}_render_c4.csv python tools/render_bench.py c4_cornell_pt --size 512 --reps 1 > gpurun_out/ncu_${tag}_r1.log 2>&1
    ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_resolve_slots|k_level_resolve|k_raygen|k_write_rgb|k_own_pixels' -s 14 -c 14 -f -o gpurun_out/prof_${tag}_render python tools/render_bench.py c4_cornell_pt --size 512 --reps 1 > gpurun_out/ncu_${tag}_r2.log 2>&1
    ncu -i gpurun_out/prof_${tag}_render.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_render_raw.csv 2>/dev/null; rm -f gpurun_out/prof_${tag}_render.ncu-rep
    ncu --set full --clock-control none -k regex:'k_trace' -s 11 -c 6 -f -o gpurun_out/prof_${tag}_render_trace python tools/render_bench.py c4_cornell_pt --size 512 --reps 1 > gpurun_out/ncu_${tag}_r3.log 2>&1
    ncu -i gpurun_out/prof_${tag}_render_trace.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_render_trace_raw.csv 2>/dev/null; rm -f gpurun_out/prof_${tag}_render_trace.ncu-rep;;
  esac
done
ls -la gpurun_out | tail -30
