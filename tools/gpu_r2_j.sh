#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/prof_session_fixed.py > gpurun_out/r2j_fixed_12m.log 2>&1; cat gpurun_out/r2j_fixed_12m.log
timeout 200 python tools/prof_session_fixed.py --n 100000008 > gpurun_out/r2j_fixed_100m.log 2>&1; cat gpurun_out/r2j_fixed_100m.log
