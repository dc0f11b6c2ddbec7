#!/bin/bash
# 8 GPUs, short: N = 8 / 4 / 1 session path in the driver's shape with the session timeline in the line
mkdir -p gpurun_out
run() {
  if [ "$1" = 1 ]; then timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2p_bench_n1.json 2>> gpurun_out/r2p_bench.err
  else timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29700 + $1)) bench.py --gpus $1 --steps 20 --warmup 5 --e2e-steps 3 > gpurun_out/r2p_bench_n$1.json 2>> gpurun_out/r2p_bench.err; fi
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2p_bench_n$1.json") if l.startswith("{")][-1])
    print("N=$1", "value %.1f Gpts/s" % (d["value"] / 1e9), "ms/step %.4f" % d["ms_per_step"], "sustained %.4f" % d["sustained"]["ms_per_step"], json.dumps(d["detail"]["session_timeline"]))
except Exception as e:
    print("N=$1 FAILED", e)
PY
}
run 8
run 1

run 8
grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/r2p_bench.err | tail -5
