#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -6 > gpurun_out/r2i_pytest.log; cat gpurun_out/r2i_pytest.log
timeout 200 python tools/prof_session.py --evals 200 --n 12500004 > gpurun_out/r2i_session_12m.log 2>&1; cat gpurun_out/r2i_session_12m.log
timeout 200 python tools/prof_session.py --evals 200 > gpurun_out/r2i_session_100m.log 2>&1; cat gpurun_out/r2i_session_100m.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2i_bench_n1_k20.json 2> gpurun_out/r2i_bench.err; tail -3 gpurun_out/r2i_bench.err; cut -c1-330 gpurun_out/r2i_bench_n1_k20.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --pts-per-room 1041667 > gpurun_out/r2i_bench_n1_k20_12m.json 2>> gpurun_out/r2i_bench.err; cut -c1-330 gpurun_out/r2i_bench_n1_k20_12m.json
#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/prof_session_fixed.py > gpurun_out/r2j_fixed_12m.log 2>&1; cat gpurun_out/r2j_fixed_12m.log
timeout 200 python tools/prof_session_fixed.py --n 100000008 > gpurun_out/r2j_fixed_100m.log 2>&1; cat gpurun_out/r2j_fixed_100m.log
