mkdir -p gpurun_out
for n in 240000 12500000 100000008; do
  timeout 200 python tools/prof_eval.py --reps 20 --var 2 --n $n --blocks 2>&1 | head -6
  timeout 200 python tools/prof_eval.py --reps 20 --n $n 2>&1 | head -1
done > gpurun_out/eval_timeline_pred.log 2>&1
cat gpurun_out/eval_timeline_pred.log
