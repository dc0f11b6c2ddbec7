#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rooms_cuboid_sums_pred -s 2 -c 1 -o gpurun_out/eval_pred -f python tools/prof_eval.py --var 2 --cons 3 --reps 1 > gpurun_out/ncu_pred.log 2>&1; tail -3 gpurun_out/ncu_pred.log
