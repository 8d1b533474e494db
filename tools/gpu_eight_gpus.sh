#!/bin/bash
# 8 GPUs: bench at N=8 (configs[3]: 1 M cubes, label all-gather; sgan leg of configs[4])
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 10 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_n8.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n8.json'))
for k in ['value','ms_per_step','e2e','parity','sgan','gpu_launches']: print(k, json.dumps(d.get(k))[:1100])"
