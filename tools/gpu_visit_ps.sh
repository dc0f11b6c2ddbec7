mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 400 python tools/bench_rows.py --only A13,A1 --out gpurun_out/rows_ps.json > gpurun_out/rows_ps.log 2>&1; cat gpurun_out/rows_ps.log
timeout 300 ncu --set full --clock-control none -k regex:"k_plane_sums_ring" -c 2 -o /tmp/rows_ncu_psr -f python tools/bench_rows.py --only A13 --reps 1 > gpurun_out/ncu_rows_psr.log 2>&1
ncu -i /tmp/rows_ncu_psr.ncu-rep --page raw --csv > gpurun_out/rows_ncu_raw_psr.csv 2>/dev/null
