#!/bin/bash
# round 2, final 1-GPU visit: the whole GPU suite, smoke, bench (both shapes) + reference arm, ncu launch list of the bench command
# (deferred-launch session: ncu blocks inside a kernel launch), one ncu --set full capture of the evaluation kernel (one-launch form)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5 > gpurun_out/r02_final_pytest_gpu.log; cat gpurun_out/r02_final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_final_smoke.log
timeout 600 python bench.py > gpurun_out/r02_final_bench_n1_k200.json 2> gpurun_out/r02_final_bench.err; tail -2 gpurun_out/r02_final_bench.err; cut -c1-300 gpurun_out/r02_final_bench_n1_k200.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench_n1_k20.json 2>> gpurun_out/r02_final_bench.err; cut -c1-300 gpurun_out/r02_final_bench_n1_k20.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_ref_n1.json 2>> gpurun_out/r02_final_bench.err; cut -c1-300 gpurun_out/r02_final_ref_n1.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --defer-launch > gpurun_out/r02_final_bench_under_ncu.log 2>&1; tail -2 gpurun_out/r02_final_bench_under_ncu.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 2 -c 1 -o /tmp/eval_r02 -f python tools/prof_eval.py --reps 1 > gpurun_out/r02_final_ncu_full.log 2>&1; tail -2 gpurun_out/r02_final_ncu_full.log
ncu -i /tmp/eval_r02.ncu-rep --page raw --csv > gpurun_out/r02_final_eval_raw.csv 2>/dev/null
ncu -i /tmp/eval_r02.ncu-rep --page details > gpurun_out/r02_final_eval_details.txt 2>/dev/null
ls -la gpurun_out/r02_final_*
