#!/bin/bash
# round 2, visit G (1 GPU): full-size tests (C4 50 M, C5 10 k frames), the other workloads at N = 1, sanitizer runs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2g_pytest_fullsize.log; cat gpurun_out/r2g_pytest_fullsize.log
for wl in c5_stream c4_cc c2_export; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r2g_${wl}_n1.json 2> gpurun_out/r2g_${wl}.err; tail -2 gpurun_out/r2g_${wl}.err; cat gpurun_out/r2g_${wl}_n1.json
done
bash tools/gpu_sanitize.sh
