#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/trace_session.py > gpurun_out/r2l_trace_12m.log 2>&1; cat gpurun_out/r2l_trace_12m.log
timeout 200 python tools/trace_session.py --n 50000004 > gpurun_out/r2l_trace_50m.log 2>&1; cat gpurun_out/r2l_trace_50m.log
