#!/bin/bash
# sanitizer pass over the round-2 kernels (k1 SUMS / direct stores, pipelined k6_tower, TMEM A operand)
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  echo "== compute-sanitizer --tool $tool tools/san_small.py"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/san_small.py > gpurun_out/sanitizer_${tool}_r3.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_${tool}_r3.log
  tail -4 gpurun_out/sanitizer_${tool}_r3.log | cut -c1-200
done
echo "== racecheck without the bulk-copy projection kernels"
timeout 900 compute-sanitizer --tool racecheck --kernel-name-exclude kns=k1_project_max --error-exitcode 9 python tools/san_small.py > gpurun_out/sanitizer_racecheck_r3.log 2>&1
echo "rc=$?" >> gpurun_out/sanitizer_racecheck_r3.log
grep -v "Saved host\|Host Frame\|=========     at" gpurun_out/sanitizer_racecheck_r3.log | tail -8 | cut -c1-250
