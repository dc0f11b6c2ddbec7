#!/bin/bash
mkdir -p gpurun_out
P="timeout 120 python tools/prof_eval.py --reps 20"
{
for c in 0 5123 4484 4486 3844 3848 0; do $P --var 7 --cons $c | tail -1; done
} 2>&1 | tee gpurun_out/sweep3.log
