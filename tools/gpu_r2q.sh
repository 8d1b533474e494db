#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py -x -q > gpurun_out/pytest_nets_r2q.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_nets_r2q.log
tail -3 gpurun_out/pytest_nets_r2q.log | cut -c1-300
grep -n "^E  " gpurun_out/pytest_nets_r2q.log | head -5 | cut -c1-200
(timeout 300 python tools/time_tower.py dnn | tail -2; timeout 300 python tools/prof_nets.py --scans 32768 --chunk 8192 --kind dnn; timeout 300 python tools/prof_nets.py --scans 8192 --chunk 4096 --kind sgan_c) > gpurun_out/time_tower_r2q.txt 2>&1
grep -v Warn gpurun_out/time_tower_r2q.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k6_tower -s 2 -c 1 -f -o gpurun_out/r2_k6_dnn python tools/prof_nets.py --scans 18944 --chunk 4736 --kind dnn > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k6_tower -s 1 -c 1 -f -o gpurun_out/r2_k6_sgan python tools/prof_nets.py --scans 8192 --chunk 4096 --kind sgan_c > /dev/null 2>&1
ls -la gpurun_out/r2_k6_*.ncu-rep
echo "== racecheck without the bulk-copy projection kernels"
timeout 900 compute-sanitizer --tool racecheck --kernel-regex-exclude kns=k1_project_max --error-exitcode 9 python tools/san_small.py > gpurun_out/sanitizer_racecheck2_r2.log 2>&1
echo "rc=$?" >> gpurun_out/sanitizer_racecheck2_r2.log
grep -v "Saved host\|Host Frame\|=========     at" gpurun_out/sanitizer_racecheck2_r2.log | tail -12 | cut -c1-250
