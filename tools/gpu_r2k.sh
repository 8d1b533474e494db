#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py -x -q > gpurun_out/pytest_nets_r2k.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_nets_r2k.log
tail -4 gpurun_out/pytest_nets_r2k.log | cut -c1-300
grep -n "^E  " gpurun_out/pytest_nets_r2k.log | head -5 | cut -c1-200
(timeout 300 python tools/time_tower.py dnn; timeout 300 python tools/prof_nets.py --scans 32768 --chunk 8192 --kind dnn) > gpurun_out/time_tower_r2k.txt 2>&1
grep -v Warn gpurun_out/time_tower_r2k.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k6_tower -s 1 -c 1 -o gpurun_out/k6_full_i python tools/bench_nets.py --scans 4096 --chunk 2048 --steps 1 > /dev/null 2>&1
ls -la gpurun_out/k6_full_i.ncu-rep
