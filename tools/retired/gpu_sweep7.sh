#!/bin/bash
mkdir -p gpurun_out
P="timeout 120 python tools/prof_eval.py --reps 20"
{
for c in 768 7682 5122 5123 7683 768; do $P --var 5 --cons $c | tail -1; done
} 2>&1 | tee gpurun_out/sweep7.log
for c in 7682 5123; do HS_MODE_3=5 HS_MODE_2=$c timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/sweep7.log; done
