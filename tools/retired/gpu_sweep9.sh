#!/bin/bash
mkdir -p gpurun_out
P="timeout 120 python tools/prof_eval.py --reps 200 --rooms 2"
{
for n in 12500000 25000000; do
for c in 38434 5133 5123 5122 7682 512 768; do $P --var 5 --cons $c --n $n | tail -1; done
$P --var 7 --n $n | tail -1
done
} 2>&1 | tee gpurun_out/sweep9.log
