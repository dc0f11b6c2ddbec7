#!/bin/bash
# round 2, visit E (8 GPUs): the driver's scaling run shape (--steps 20 --warmup 5) at N = 1, 2, 4, 8, session path; launch path at 8; reference arm at 8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv | head -3
run() { # N path tag extra
  if [ "$1" = 1 ]; then timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --path $2 $4 > gpurun_out/r2e_bench_n1_$3.json 2>> gpurun_out/r2e_bench.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + $1)) bench.py --gpus $1 --steps 20 --warmup 5 --path $2 $4 > gpurun_out/r2e_bench_n$1_$3.json 2>> gpurun_out/r2e_bench.err; fi
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2e_bench_n$1_$3.json") if l.startswith("{")][-1])
    print("N=$1 $3", "value %.1f Gpts/s" % (d["value"] / 1e9), "ms/step %.4f" % d["ms_per_step"], "sustained %.4f" % d["sustained"]["ms_per_step"], "launch-form %.4f" % d["roofline"]["one_launch_per_evaluation"]["kernel_ms"], "e2e %.2f Gpts/s" % (d["e2e"]["value"] / 1e9), "frac %.3f" % d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("N=$1 $3 FAILED", e)
PY
}
run 1 session session --no-cpu-baseline
run 2 session session
run 4 session session
run 8 session session
run 8 launch launch
run 8 session session2
run 1 session session2 --no-cpu-baseline
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --impl reference --gpus 8 --steps 5 --warmup 2 > gpurun_out/r2e_ref_n8.json 2>> gpurun_out/r2e_bench.err; cut -c1-400 gpurun_out/r2e_ref_n8.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2e_ref_n1.json 2>> gpurun_out/r2e_bench.err; cut -c1-400 gpurun_out/r2e_ref_n1.json
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/r2e_bench.err | tail -15
