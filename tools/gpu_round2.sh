#!/bin/bash
# GPU-box visit for the secondary rows: parity tests (all failures listed), per-row timings in both kernel forms,
# and full-set ncu captures of the new kernels (raw CSV only).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_rows.py --out gpurun_out/rows.json > gpurun_out/rows.log 2>&1; cat gpurun_out/rows.log
cap() {  # cap <tag> <rows> <kernel regex> <count>
  timeout 400 ncu --set full --clock-control none -k regex:"$3" -c $4 -o /tmp/rows_ncu_$1 -f python tools/bench_rows.py --n 24000000 --only $2 --reps 1 --frames 100 > gpurun_out/ncu_rows_$1.log 2>&1; tail -1 gpurun_out/ncu_rows_$1.log
  ncu -i /tmp/rows_ncu_$1.ncu-rep --page raw --csv > gpurun_out/rows_ncu_raw_$1.csv 2>/dev/null
}
cap bp A1 "k_bp_" 4
cap ne A1 "k_reduce6x6_f32" 2
cap ps A13 "k_plane_sums_f32" 2
cap sel A12 "k_sel2|k_filter" 8
cap pa A5 "k_plane_assign" 2
ls -la gpurun_out/*.csv
