#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "rooms or eval or session or bfgs or smoke" 2>&1 | tail -3
for sc in 768 1536 2304 3072; do
  echo "seg_cost $sc"; HS_MODE_2=$sc timeout 100 python tools/trace_session.py 2>&1 | tail -3 | cut -c1-420; timeout 100 python tools/prof_session.py --evals 200 --n 12500004 --seg-cost $sc 2>&1 | grep -E "commit-to-commit|bit-identical"
done > gpurun_out/r2n_segcost.log 2>&1; cat gpurun_out/r2n_segcost.log
