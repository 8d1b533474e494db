#!/bin/bash
# final-tree checks: the full-size network property test, then memcheck / synccheck over tools/san_small.py
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_fullsize_r4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_fullsize_r4.log
tail -3 gpurun_out/pytest_fullsize_r4.log | cut -c1-300
grep -n "^E  " gpurun_out/pytest_fullsize_r4.log | head -8 | cut -c1-250
for tool in memcheck synccheck; do
  echo "== compute-sanitizer --tool $tool tools/san_small.py"
  timeout 240 compute-sanitizer --tool $tool --error-exitcode 9 python tools/san_small.py > gpurun_out/sanitizer_${tool}_r4.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_${tool}_r4.log
  tail -4 gpurun_out/sanitizer_${tool}_r4.log | cut -c1-200
done
