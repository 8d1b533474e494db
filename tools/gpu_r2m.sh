#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py -x -q > gpurun_out/pytest_nets_r2m.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_nets_r2m.log
tail -3 gpurun_out/pytest_nets_r2m.log | cut -c1-300
grep -n "^E  " gpurun_out/pytest_nets_r2m.log | head -5 | cut -c1-200
(timeout 300 python tools/prof_nets.py --scans 8192 --chunk 4096 --kind sgan_c) > gpurun_out/time_tower_r2m.txt 2>&1
grep -v Warn gpurun_out/time_tower_r2m.txt
