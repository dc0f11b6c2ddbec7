#!/bin/bash
# 2 GPUs: validation after the launch-order / ctl changes + timeline stats in the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -4 > gpurun_out/r2o_pytest.log; cat gpurun_out/r2o_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2o_bench_n2.json 2>> gpurun_out/r2o_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29912 bench.py --gpus 2 --steps 20 --warmup 5 --pts-per-room 2083334 > gpurun_out/r2o_bench_n2_small.json 2>> gpurun_out/r2o_bench.err
python - <<PY
import json
for f in ("n1", "n2", "n2_small"):
    d = json.loads([l for l in open(f"gpurun_out/r2o_bench_{f}.json") if l.startswith("{")][-1])
    print(f, "ms/step %.4f" % d["ms_per_step"], "sustained %.4f" % d["sustained"]["ms_per_step"], json.dumps(d["detail"]["session_timeline"]))
PY
grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/r2o_bench.err | tail -5
