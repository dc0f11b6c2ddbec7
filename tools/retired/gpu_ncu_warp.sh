#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rooms_cuboid_sums_warp -s 2 -c 1 -o gpurun_out/eval_warp -f python tools/prof_eval.py --var 5 --cons ${CONS:-5123} --reps 1 > gpurun_out/ncu_warp.log 2>&1; tail -2 gpurun_out/ncu_warp.log
