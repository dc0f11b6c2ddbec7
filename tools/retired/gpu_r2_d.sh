#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sweep_eval.py --vars 0,1,2,3,5 > gpurun_out/r2d2_sweep_100m.log 2>&1; cat gpurun_out/r2d2_sweep_100m.log
timeout 300 python tools/sweep_eval.py --n 12500004 --reps 300 --vars 0,1,2,3,5 > gpurun_out/r2d2_sweep_12m.log 2>&1; cat gpurun_out/r2d2_sweep_12m.log
