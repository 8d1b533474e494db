#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r3d.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r3d.log
tail -3 gpurun_out/pytest_gpu_r3d.log | cut -c1-300
grep -n "^E  " gpurun_out/pytest_gpu_r3d.log | head -5 | cut -c1-200
(timeout 300 python tools/prof_nets.py --scans 32768 --chunk 8192 --kind dnn; timeout 300 python tools/prof_nets.py --scans 8192 --chunk 4096 --kind sgan_c) > gpurun_out/time_tower_r3d.txt 2>&1
grep -v Warn gpurun_out/time_tower_r3d.txt
timeout 600 python bench.py --steps 10 > gpurun_out/bench_r3d.json 2> gpurun_out/bench_r3d.err
echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r3d.json'))
for k in ['value','dnn','sgan','general_precision']: print(k, json.dumps(d.get(k))[:300])"
