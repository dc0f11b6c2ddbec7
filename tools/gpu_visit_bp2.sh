mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 400 python tools/bench_rows.py --only A1 --out gpurun_out/rows_bp2.json > gpurun_out/rows_bp2.log 2>&1; cat gpurun_out/rows_bp2.log
timeout 300 ncu --set full --clock-control none -k regex:"k_bp_onepass|k_reduce6x6_f32" -c 3 -o /tmp/rows_ncu_bp2 -f python tools/bench_rows.py --n 24000000 --only A1 --reps 1 --frames 100 > gpurun_out/ncu_rows_bp2.log 2>&1
ncu -i /tmp/rows_ncu_bp2.ncu-rep --page raw --csv > gpurun_out/rows_ncu_raw_bp2.csv 2>/dev/null
