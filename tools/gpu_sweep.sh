#!/bin/bash
# parity first, then time the evaluation-kernel variants (never a bench number: tools/prof_eval.py)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for c in 0 1 2 3 4 5; do timeout 120 python tools/prof_eval.py --cons $c --reps 20 2>&1 | tail -1; done
timeout 120 python tools/prof_eval.py --var 3 --reps 20 2>&1 | tail -1
