#!/bin/bash
mkdir -p gpurun_out
P="timeout 120 python tools/prof_eval.py --reps 20"
HS_MODE_3=6 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/sweep6.log
{
$P --var 7 | tail -1
for c in 0 1614 1618 1623 2414 2416 3214; do $P --var 6 --cons $c | tail -1; done
} 2>&1 | tee -a gpurun_out/sweep6.log
HS_MODE_3=6 HS_MODE_2=2414 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/sweep6.log
