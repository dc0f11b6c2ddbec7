mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_rows.py --only A1,A12 --out gpurun_out/rows_a1.json > gpurun_out/rows_a1.log 2>&1; cat gpurun_out/rows_a1.log
timeout 400 ncu --set full --clock-control none -k regex:"k_bp_" -c 4 -o /tmp/rows_ncu_bp -f python tools/bench_rows.py --n 24000000 --only A1 --reps 1 --frames 100 > gpurun_out/ncu_rows_bp.log 2>&1
ncu -i /tmp/rows_ncu_bp.ncu-rep --page raw --csv > gpurun_out/rows_ncu_raw_bp.csv 2>/dev/null
