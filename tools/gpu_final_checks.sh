#!/bin/bash
# final-tree checks: the network tests (incl. old-form / new-form agreement), memcheck / synccheck over
# tools/san_small.py, sgan chunk sweep
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_nets.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_nets_r4e.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_nets_r4e.log
tail -3 gpurun_out/pytest_nets_r4e.log | cut -c1-300
grep -n "^E  " gpurun_out/pytest_nets_r4e.log | head -8 | cut -c1-250
for tool in memcheck synccheck; do
  echo "== compute-sanitizer --tool $tool tools/san_small.py"
  timeout 240 compute-sanitizer --tool $tool --error-exitcode 9 python tools/san_small.py > gpurun_out/sanitizer_${tool}_r4e.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_${tool}_r4e.log
  tail -4 gpurun_out/sanitizer_${tool}_r4e.log | cut -c1-200
done
for ch in 2048 8192; do
  timeout 200 python tools/prof_nets.py --scans 16384 --chunk $ch --kind sgan_c 2>&1 | grep "ms/pass\|sum of"
done
