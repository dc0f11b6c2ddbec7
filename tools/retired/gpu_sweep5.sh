#!/bin/bash
mkdir -p gpurun_out
P="timeout 120 python tools/prof_eval.py --reps 20"
{
$P --var 7 | tail -1
for c in 512 640 768 896 992; do $P --var 5 --cons $c | tail -1; done
} 2>&1 | tee gpurun_out/sweep5.log
HS_MODE_3=5 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee -a gpurun_out/sweep5.log
