#!/bin/bash
mkdir -p gpurun_out
P="timeout 120 python tools/prof_eval.py --reps 20"
{
for c in ${CONS:-5133 38434 51224 48033 25636 5133}; do $P --var 5 --cons $c | tail -1; done
} 2>&1 | tee gpurun_out/sweep8.log
