#!/bin/bash
# round 2, visit F (1 GPU): full GPU suite (BFGS parity, sharded export, peer tests), bench in the driver's shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -25 > gpurun_out/r2f_pytest.log; cat gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_n1_k20.json 2> gpurun_out/r2f_bench.err; tail -3 gpurun_out/r2f_bench.err; cat gpurun_out/r2f_bench_n1_k20.json
timeout 600 python bench.py > gpurun_out/r2f_bench_n1.json 2>> gpurun_out/r2f_bench.err; cat gpurun_out/r2f_bench_n1.json
