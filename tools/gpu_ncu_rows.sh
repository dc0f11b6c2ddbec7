#!/bin/bash
# full-set ncu capture of the secondary kernels; only the raw CSV comes back (the report itself is too big for gpurun_out)
mkdir -p gpurun_out
for spec in ${SPECS:-"A1:9" "A10:6" "A12:12"}; do
  rows=${spec%%:*}; cnt=${spec##*:}
  timeout 400 ncu --set full --clock-control none -k regex:"k_reduce6x6|k_plane_sums|k_bp_|k_cc_|k_sel_pass|k_filter|k_plane_assign|k_affine" -c $cnt -o /tmp/rows_ncu_$rows -f python tools/bench_rows.py --n 24000000 --only $rows --reps 1 --frames 100 --cc-planes 8 > gpurun_out/ncu_rows_$rows.log 2>&1; tail -2 gpurun_out/ncu_rows_$rows.log
  ncu -i /tmp/rows_ncu_$rows.ncu-rep --page raw --csv > gpurun_out/rows_ncu_raw_$rows.csv 2>/dev/null
done
ls -la gpurun_out/*.csv
