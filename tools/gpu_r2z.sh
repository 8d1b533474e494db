#!/bin/bash
tag=${1:-r2z}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py -x -q > gpurun_out/pytest_nets_$tag.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_nets_$tag.log
tail -3 gpurun_out/pytest_nets_$tag.log | cut -c1-300
(echo "early gather (default):"; timeout 300 python tools/time_tower.py dnn | tail -2
 echo "gather after the commit (RML_T6_DBG=32):"; RML_T6_DBG=32 timeout 300 python tools/time_tower.py dnn | tail -2
 timeout 300 python tools/prof_nets.py --scans 32768 --chunk 8192 --kind dnn) > gpurun_out/time_tower_$tag.txt 2>&1
grep -v Warn gpurun_out/time_tower_$tag.txt
