#!/bin/bash
# round 2, visit A: parity tests with the rewritten evaluation kernel, session timing, bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -25 > gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_pytest.log
timeout 200 python tools/prof_session.py --evals 200 > gpurun_out/r2a_session_100m.log 2>&1; cat gpurun_out/r2a_session_100m.log
timeout 200 python tools/prof_session.py --evals 200 --n 12500004 > gpurun_out/r2a_session_12m.log 2>&1; cat gpurun_out/r2a_session_12m.log
timeout 200 python tools/prof_session.py --evals 2000 --n 12500004 > gpurun_out/r2a_session_12m_2000.log 2>&1; cat gpurun_out/r2a_session_12m_2000.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -3 gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json
