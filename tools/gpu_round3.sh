#!/bin/bash
# Round-end GPU-box visit: parity tests, bench line (+ reference arm), ncu launch list of the bench command, one full capture of
# the headline kernel, per-row timings, per-row ncu captures, per-block timeline of the headline kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rooms_cuboid_sums -s 2 -c 1 -o /tmp/eval_default -f python tools/prof_eval.py --reps 1 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ncu -i /tmp/eval_default.ncu-rep --page raw --csv > gpurun_out/eval_default_raw.csv 2>/dev/null
ncu -i /tmp/eval_default.ncu-rep --page details > gpurun_out/eval_default_details.txt 2>/dev/null
timeout 300 python tools/prof_eval.py --reps 20 > gpurun_out/eval_isolated.log 2>&1; cat gpurun_out/eval_isolated.log
timeout 600 python tools/bench_rows.py --out gpurun_out/rows.json > gpurun_out/rows.log 2>&1; cat gpurun_out/rows.log
cap() {  # cap <tag> <rows> <kernel regex> <count>
  timeout 400 ncu --set full --clock-control none -k regex:"$3" -c $4 -o /tmp/rows_ncu_$1 -f python tools/bench_rows.py --only $2 --reps 1 --frames 200 > gpurun_out/ncu_rows_$1.log 2>&1; tail -1 gpurun_out/ncu_rows_$1.log
  ncu -i /tmp/rows_ncu_$1.ncu-rep --page raw --csv > gpurun_out/rows_ncu_raw_$1.csv 2>/dev/null
}
cap bp A1 "k_bp_onepass" 2
cap ne A1 "k_reduce6x6_f32" 2
cap ps A13 "k_plane_sums_ring" 2
cap sel A12 "k_sel2" 4
cap flt A12 "k_filter" 4
cap pa A5 "k_plane_assign" 2
ls -la gpurun_out/
