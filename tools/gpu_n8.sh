#!/bin/bash
# 8-GPU visit: strong scaling of the 100 M-point apartment at N = 8 (peer-memory all-reduce and NCCL), N = 4, and weak scaling at 8
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 200 --warmup 10 ${@:3} 2>&1 | tail -1 > gpurun_out/$2.json; cut -c1-200 gpurun_out/$2.json; }
run 8 bench_n8_p2p --collective p2p
run 8 bench_n8_nccl --collective nccl
run 4 bench_n4_p2p --collective p2p
run 8 bench_n8_weak_p2p --collective p2p --scaling weak
