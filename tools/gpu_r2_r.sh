#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -4 | tee gpurun_out/r02_final_pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_final_smoke.log
