#!/bin/bash
# 2 GPUs: peer + session tests after widening the window, N=1/N=2 bench, single-GPU timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -6 > gpurun_out/r2m_pytest.log; cat gpurun_out/r2m_pytest.log
timeout 200 python tools/trace_session.py > gpurun_out/r2m_trace_12m.log 2>&1; tail -4 gpurun_out/r2m_trace_12m.log
timeout 200 python tools/prof_session_fixed.py > gpurun_out/r2m_fixed_12m.log 2>&1; tail -2 gpurun_out/r2m_fixed_12m.log
timeout 100 python tools/prof_session.py --evals 200 --n 12500004 2>&1 | grep -E "commit-to-commit|raw C|bit-identical" | tee gpurun_out/r2m_session_12m.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench.err; cut -c1-330 gpurun_out/r2m_bench_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2m_bench_n2.json 2>> gpurun_out/r2m_bench.err; cut -c1-330 gpurun_out/r2m_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus 2 --steps 20 --warmup 5 --pts-per-room 2083334 > gpurun_out/r2m_bench_n2_small.json 2>> gpurun_out/r2m_bench.err; cut -c1-330 gpurun_out/r2m_bench_n2_small.json
grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/r2m_bench.err | tail -5
