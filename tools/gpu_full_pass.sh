#!/bin/bash
# round-end style pass on one B200: smoke, every -m gpu test, the N=1 bench (both arms)
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_full.log
tail -4 gpurun_out/pytest_gpu_full.log | cut -c1-300
grep -n "^E  " gpurun_out/pytest_gpu_full.log | head -5 | cut -c1-200
timeout 900 python bench.py --steps 10 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_full.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full.json'))
for k in ['value','roofline','e2e','parity','dnn','sgan','general_precision','latency']: print(k, json.dumps(d.get(k))[:700])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_full_ref.json 2> gpurun_out/bench_full_ref.err
echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_full_ref.json
