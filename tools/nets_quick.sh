timeout 300 python -m pytest tests/test_gpu_nets.py -x -q 2>&1 | tail -3 | tee /tmp/nq.log
grep -q "failed\|error" /tmp/nq.log && exit 1
for ch in 192 512 1024; do timeout 120 python tools/bench_nets.py --scans 16384 --chunk $ch --steps 2 2>&1 | tail -2; done
RML_NET_FUSE1=0 timeout 120 python tools/bench_nets.py --scans 16384 --chunk 1024 --steps 2 2>&1 | tail -2
