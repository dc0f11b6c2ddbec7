#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 90 -x -k "nm_fit or bfgs or smoke or eval_random" 2>&1 | tail -6 | tee gpurun_out/r2q_pytest.log
