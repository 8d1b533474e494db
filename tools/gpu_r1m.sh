# Round-1 GPU pass M: dense stack with 4 K blocks per ring stage
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nets.py tests/test_gpu_u8cubes.py -q -p no:cacheprovider -k "nets or net_" --timeout 200 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_nets.log
echo "== default (kpg 4)"; timeout 200 python tools/bench_nets.py --scans 16384 --chunk 1024 --steps 3 2>&1 | tail -2 | tee gpurun_out/nets_m.txt
echo "== RML_K5_KPG=1"; RML_K5_KPG=1 timeout 200 python tools/bench_nets.py --scans 16384 --chunk 1024 --steps 3 2>&1 | tail -2 | tee -a gpurun_out/nets_m.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k5_" -c 4 --csv --log-file gpurun_out/launches_k5_m.csv python tools/bench_nets.py --scans 4096 --chunk 1024 --steps 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_k5_m.csv | cut -d'"' -f10,18,30 | tail -4
