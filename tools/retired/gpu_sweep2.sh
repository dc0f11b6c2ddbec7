#!/bin/bash
# variant sweep of the cuboid-sums kernel (never a bench number: tools/prof_eval.py), then parity of the candidates
mkdir -p gpurun_out
P="timeout 120 python tools/prof_eval.py --reps 20"
{
$P --var 7 | tail -1
$P --var 7 --cons 256 --bps 2 | tail -1
$P --var 7 --cons 160 --bps 3 | tail -1
$P --var 4 | tail -1
$P --var 4 --cons 480 | tail -1
$P --var 4 --cons 256 --bps 2 | tail -1
$P --var 7 | tail -1
} 2>&1 | tee gpurun_out/sweep2.log
HS_MODE_3=4 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee -a gpurun_out/sweep2.log
HS_MODE_3=4 HS_MODE_2=256 HS_MODE_1=2 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee -a gpurun_out/sweep2.log
