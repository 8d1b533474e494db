#!/bin/bash
# every -m gpu test, then the sgan and dnn per-kernel splits
tag=${1:-r4c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log
tail -3 gpurun_out/pytest_gpu_$tag.log | cut -c1-300
grep -n "^E  " gpurun_out/pytest_gpu_$tag.log | head -5 | cut -c1-200
(timeout 300 python tools/prof_nets.py --scans 8192 --chunk 4096 --kind sgan_c
 timeout 300 python tools/prof_nets.py --scans 32768 --chunk 8192 --kind dnn) > gpurun_out/time_nets_$tag.txt 2>&1
grep -v "Warn\|_warn" gpurun_out/time_nets_$tag.txt
