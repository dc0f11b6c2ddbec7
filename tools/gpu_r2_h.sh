#!/bin/bash
# round 2, visit H (2 GPUs): peer tests after the exchange rewrite, bench N = 2 (session / launch), workloads at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -q --timeout 300 -x 2>&1 | tail -5 > gpurun_out/r2h_pytest_peer.log; cat gpurun_out/r2h_pytest_peer.log
for path in session launch; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 --path $path > gpurun_out/r2h_bench_n2_$path.json 2>> gpurun_out/r2h_bench.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2h_bench_n2_$path.json") if l.startswith("{")][-1])
print("N=2 $path", "value %.1f Gpts/s" % (d["value"] / 1e9), "ms/step %.4f" % d["ms_per_step"], "sustained %.4f" % d["sustained"]["ms_per_step"], "launch-form %.4f" % d["roofline"]["one_launch_per_evaluation"]["kernel_ms"], d["detail"]["collective_vs_nccl_max_rel"])
PY
done
for wl in c5_stream c4_cc c2_export a12_ceiling; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --workload $wl --steps 10 --warmup 3 > gpurun_out/r2h_${wl}_n2.json 2> gpurun_out/r2h_${wl}.err; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/r2h_${wl}.err | tail -3; cut -c1-700 gpurun_out/r2h_${wl}_n2.json
done
timeout 300 python bench.py --workload a12_ceiling --steps 10 --warmup 3 > gpurun_out/r2h_a12_ceiling_n1.json 2> gpurun_out/r2h_a12.err; tail -2 gpurun_out/r2h_a12.err; cut -c1-700 gpurun_out/r2h_a12_ceiling_n1.json
