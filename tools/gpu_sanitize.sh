#!/bin/bash
# compute-sanitizer over the lock-free / barrier-heavy kernels: memcheck + synccheck on the GPU tests that hit
# union-find, decoupled look-back, ticket / last-block reductions, the evaluation kernel (one-launch and session forms, mbarrier
# ring) and the peer exchange; racecheck (shared-memory hazards) on the kernels that synchronise with barriers only.  The evaluation
# kernel's warps talk through volatile shared-memory queues and mbarriers by design, which racecheck cannot model: it is excluded
# from racecheck by kernel name and covered by memcheck / synccheck / the bit-identity tests instead.
mkdir -p gpurun_out
SEL='cc or filter or kth or remove_ceiling or backproject_ref or plane_assign or plane_sums or scatter or mean_extent or ragged or additive'
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck; do
  timeout 1200 $CS --tool $tool --error-exitcode 9 --launch-timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -k "$SEL or eval_random or session" > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/r02_sanitizer_$tool.log
done
timeout 1200 $CS --tool racecheck --error-exitcode 9 --kernel-name-exclude kns=k_eval --kernel-name-exclude kns=k_peer python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -k "$SEL" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_racecheck.log
