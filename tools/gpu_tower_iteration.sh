#!/bin/bash
# tower iteration (short): parity, tower timing, per-kernel split dnn + sgan, one ncu capture
tag=${1:-r2x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py -x -q > gpurun_out/pytest_nets_$tag.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_nets_$tag.log
tail -3 gpurun_out/pytest_nets_$tag.log | cut -c1-300
grep -n "^E  " gpurun_out/pytest_nets_$tag.log | head -5 | cut -c1-200
(timeout 300 python tools/time_tower.py dnn | tail -2
 timeout 300 python tools/prof_nets.py --scans 32768 --chunk 8192 --kind dnn
 timeout 300 python tools/prof_nets.py --scans 8192 --chunk 4096 --kind sgan_c) > gpurun_out/time_tower_$tag.txt 2>&1
grep -v Warn gpurun_out/time_tower_$tag.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k6_tower -s 2 -c 1 -f -o gpurun_out/${tag}_k6_dnn python tools/prof_nets.py --scans 18944 --chunk 4736 --kind dnn > /dev/null 2>&1
ls -la gpurun_out/${tag}_k6_dnn.ncu-rep
