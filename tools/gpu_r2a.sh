#!/bin/bash
# round 2, call A: probe (elect issue cost), full GPU test-suite, 1-GPU bench
mkdir -p gpurun_out
./tools/bin/umma_probe > gpurun_out/umma_probe2.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2a.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2a.log
tail -5 gpurun_out/pytest_gpu_r2a.log
timeout 900 python bench.py --steps 10 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
echo "bench rc=$?"
tail -3 gpurun_out/bench_r2a.err
cat gpurun_out/bench_r2a.json | head -c 6000
