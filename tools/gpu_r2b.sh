#!/bin/bash
# round 2, call B: fused tower kernel — parity, throughput, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nets.py -x -q > gpurun_out/pytest_nets_r2b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_nets_r2b.log
tail -25 gpurun_out/pytest_nets_r2b.log
timeout 600 python tools/bench_nets.py --scans 16384 --chunk 4096 > gpurun_out/nets_r2b.txt 2>&1
cat gpurun_out/nets_r2b.txt
RML_NET_TOWER=0 timeout 600 python tools/bench_nets.py --scans 16384 --chunk 4096 > gpurun_out/nets_r2b_old.txt 2>&1
cat gpurun_out/nets_r2b_old.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_nets_r2b.csv python tools/bench_nets.py --scans 4096 --chunk 2048 --steps 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_nets_r2b.csv | cut -d, -f5,8,9,15 | tail -30
