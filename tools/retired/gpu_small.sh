#!/bin/bash
mkdir -p gpurun_out
{
for c in 38434 38435; do for n in 240000 12500000 100000008; do timeout 120 python tools/prof_eval.py --var 5 --cons $c --reps 200 --rooms 2 --n $n | tail -1; done; done
} 2>&1 | tee gpurun_out/small.log
