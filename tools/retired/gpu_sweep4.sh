#!/bin/bash
mkdir -p gpurun_out
P="timeout 120 python tools/prof_eval.py --reps 20"
{
for ps in 0 20 50 100 200 400 1000 0; do $P --var 7 --psleep $ps | tail -1; done
$P --var 7 --cons 3844 --psleep 100 | tail -1
$P --var 4 --psleep 100 | tail -1
} 2>&1 | tee gpurun_out/sweep4.log
