#!/bin/bash
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 80 -x -k "plane_sums or pipeline or unpaired" 2>&1 | tail -4 | tee gpurun_out/r2s_pytest.log
timeout 90 python tools/bench_rows.py --only A13 --out gpurun_out/r2s_rows_a13.json 2>&1 | tail -5 | tee gpurun_out/r2s_rows_a13.log
