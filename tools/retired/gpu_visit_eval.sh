mkdir -p gpurun_out
for n in 240000 12500000 100000008; do timeout 200 python tools/prof_eval.py --reps 20 --n $n 2>&1 | head -1; done > gpurun_out/eval_after.log 2>&1
timeout 200 python tools/prof_eval.py --reps 20 --n 12500000 --blocks2 2>&1 | sed -n 2,5p >> gpurun_out/eval_after.log
cat gpurun_out/eval_after.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -k "cuboid or rooms or smoke or apartment" --timeout 200 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])"
