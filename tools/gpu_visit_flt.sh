mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 400 python tools/bench_rows.py --only A12,A1 --out gpurun_out/rows_flt.json > gpurun_out/rows_flt.log 2>&1; cat gpurun_out/rows_flt.log
