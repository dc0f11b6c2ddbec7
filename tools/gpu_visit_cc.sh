mkdir -p gpurun_out
for c in 1 2 4 8 16; do echo "chunks=$c"; HS_MODE_11=$c timeout 300 python tools/bench_rows.py --only A10 --out gpurun_out/rows_cc$c.json 2>&1 | tail -1 | cut -c1-200; done > gpurun_out/rows_cc.log 2>&1; cat gpurun_out/rows_cc.log
