#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2l.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2l.log
tail -4 gpurun_out/pytest_gpu_r2l.log | cut -c1-300
grep -n "^E  " gpurun_out/pytest_gpu_r2l.log | head -5 | cut -c1-200
(timeout 300 python tools/time_tower.py dnn | tail -2; timeout 300 python tools/prof_nets.py --scans 32768 --chunk 8192 --kind dnn; timeout 300 python tools/prof_nets.py --scans 8192 --chunk 4096 --kind sgan_c) > gpurun_out/time_tower_r2l.txt 2>&1
grep -v Warn gpurun_out/time_tower_r2l.txt
timeout 900 python bench.py --steps 10 > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_r2l.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2l.json'))
for k in ['value','e2e','dnn','sgan','general_precision']: print(k, json.dumps(d.get(k))[:600])"
