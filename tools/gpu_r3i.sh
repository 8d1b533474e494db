#!/bin/bash
mkdir -p gpurun_out
nproc; lscpu | grep -i "model name" | head -1
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -k "narrowing or predict_targets or real_valued" 2>&1 | tail -4
timeout 900 python bench.py --steps 10 > gpurun_out/bench_r3i.json 2> gpurun_out/bench_r3i.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_r3i.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r3i.json'))
for k in ['value','e2e','latency','derived_targets']: print(k, json.dumps(d.get(k))[:1500])"
