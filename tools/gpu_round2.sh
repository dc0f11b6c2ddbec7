#!/bin/bash
# GPU-box visit for the secondary rows: parity tests (all failures listed), per-row timings in both kernel forms,
# and a full-set ncu capture of the new kernels (raw CSV only).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_rows.py --out gpurun_out/rows.json > gpurun_out/rows.log 2>&1; cat gpurun_out/rows.log
for spec in "A1:8" "A13:3" "A12:10" "A5:4"; do
  rows=${spec%%:*}; cnt=${spec##*:}
  timeout 400 ncu --set full --clock-control none -k regex:"k_reduce6x6_f32|k_plane_sums_f32|k_bp_|k_sel2|k_filter|k_plane_assign" -c $cnt -o /tmp/rows_ncu_$rows -f python tools/bench_rows.py --n 24000000 --only $rows --reps 1 --frames 100 > gpurun_out/ncu_rows_$rows.log 2>&1; tail -2 gpurun_out/ncu_rows_$rows.log
  ncu -i /tmp/rows_ncu_$rows.ncu-rep --page raw --csv > gpurun_out/rows_ncu_raw_$rows.csv 2>/dev/null
done
ls -la gpurun_out/*.csv
