mkdir -p gpurun_out
for n in 240000 12500000 100000008; do
  timeout 200 python tools/prof_eval.py --reps 20 --n $n --blocks2 2>&1 | head -6
done > gpurun_out/eval_timeline_warp.log 2>&1
cat gpurun_out/eval_timeline_warp.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "cuboid or rooms or smoke" --timeout 120 2>&1 | tail -2
