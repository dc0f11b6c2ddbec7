#!/bin/bash
mkdir -p gpurun_out
for sc in 768 1536 2304 3072 4608 6144; do
  echo "seg_cost $sc"; timeout 100 python tools/prof_session.py --evals 200 --n 12500004 --seg-cost $sc 2>&1 | grep -E "commit-to-commit|bit-identical"
done > gpurun_out/r2k_segcost_12m.log 2>&1; cat gpurun_out/r2k_segcost_12m.log
for n in 25000008 50000004; do echo "n $n"; timeout 100 python tools/prof_session.py --evals 200 --n $n 2>&1 | grep -E "commit-to-commit"; done > gpurun_out/r2k_sizes.log 2>&1; cat gpurun_out/r2k_sizes.log
