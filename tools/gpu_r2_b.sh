#!/bin/bash
# round 2, visit B (2 GPUs): peer tests, session timing, bench at N=1 and N=2 (session and launch paths)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_parity.py -m gpu -q --timeout 300 -x 2>&1 | tail -25 > gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_pytest.log
timeout 200 python tools/prof_session.py --evals 200 --n 12500004 > gpurun_out/r2b_session_12m.log 2>&1; cat gpurun_out/r2b_session_12m.log
timeout 200 python tools/prof_session.py --evals 200 > gpurun_out/r2b_session_100m.log 2>&1; cat gpurun_out/r2b_session_100m.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench.err; tail -3 gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench_n1.json
timeout 600 python bench.py --steps 20 --warmup 5 --path launch --no-cpu-baseline > gpurun_out/r2b_bench_n1_launch.json 2>> gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench_n1_launch.json
for path in session launch; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --path $path > gpurun_out/r2b_bench_n2_$path.json 2>> gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench_n2_$path.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r2b_bench_n2_200.json 2>> gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench_n2_200.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2b_ref_n1.json 2>> gpurun_out/r2b_bench.err; cat gpurun_out/r2b_ref_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/r2b_ref_n2.json 2>> gpurun_out/r2b_bench.err; cat gpurun_out/r2b_ref_n2.json
tail -20 gpurun_out/r2b_bench.err
