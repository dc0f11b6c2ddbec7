#!/bin/bash
# round 2, visit C (1 GPU): whole GPU suite with the new point block, session timing, bench, ncu of the one-launch kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -25 > gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_pytest.log
timeout 200 python tools/prof_session.py --evals 200 > gpurun_out/r2c_session_100m.log 2>&1; cat gpurun_out/r2c_session_100m.log
timeout 200 python tools/prof_session.py --evals 200 --n 12500004 > gpurun_out/r2c_session_12m.log 2>&1; cat gpurun_out/r2c_session_12m.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_n1_k20.json 2> gpurun_out/r2c_bench.err; tail -3 gpurun_out/r2c_bench.err; cat gpurun_out/r2c_bench_n1_k20.json
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c_bench_n1.json 2>> gpurun_out/r2c_bench.err; cat gpurun_out/r2c_bench_n1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 2 -c 1 -o /tmp/eval_r2c -f python tools/prof_eval.py --reps 1 > gpurun_out/r2c_ncu_full.log 2>&1; tail -2 gpurun_out/r2c_ncu_full.log
ncu -i /tmp/eval_r2c.ncu-rep --page raw --csv > gpurun_out/r2c_eval_raw.csv 2>/dev/null
ncu -i /tmp/eval_r2c.ncu-rep --page details > gpurun_out/r2c_eval_details.txt 2>/dev/null
