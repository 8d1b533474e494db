#!/bin/bash
# 2 GPUs: C-ABI all-gather test, distributed determinism test, bench at N=2 (both arms)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n2.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_round2.py -x -q -k "dist or allgather or sharded" > gpurun_out/pytest_dist_n2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_dist_n2.log
tail -5 gpurun_out/pytest_dist_n2.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 > gpurun_out/bench_n2_n2.json 2> gpurun_out/bench_n2_n2.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_n2_n2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_n2.json'))
for k in ['value','ms_per_step','e2e','parity','sgan','gpu_launches']: print(k, json.dumps(d.get(k))[:900])"
