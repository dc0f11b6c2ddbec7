#!/bin/bash
# parity first, then time the evaluation-kernel variants (never a bench number: tools/prof_eval.py)
mkdir -p gpurun_out
for c in ${CONS:-0 1 2 3 4 5 6 7 8}; do timeout 120 python tools/prof_eval.py --var 2 --cons $c --reps 20 2>&1 | tail -1; done
timeout 120 python tools/prof_eval.py --var 7 --reps 20 2>&1 | tail -1
HS_MODE_3=2 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
