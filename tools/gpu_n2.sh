#!/bin/bash
# 2-GPU visit: peer-memory all-reduce vs NCCL, both through bench.py (strong scaling, 100 M points total)
mkdir -p gpurun_out
N=${N:-2}
for coll in p2p nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 --collective $coll 2>&1 | tail -2 | tee gpurun_out/bench_n${N}_$coll.json
done
