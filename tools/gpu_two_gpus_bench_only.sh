mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 > gpurun_out/bench_r4g_n2.json 2> gpurun_out/bench_r4g_n2.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_r4g_n2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r4g_n2.json'))
for k in ['value','ms_per_step','e2e','parity','sgan','gpu_launches']: print(k, json.dumps(d.get(k))[:1100])"
