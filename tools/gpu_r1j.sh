# Round-1 GPU pass J: implicit-GEMM layers with branch-resident weights
mkdir -p gpurun_out
date +%T
timeout 400 python -m pytest tests/test_gpu_nets.py tests/test_gpu_u8cubes.py -q -p no:cacheprovider -k "nets or net_" --timeout 200 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_nets.log
date +%T
echo "== default (resident weights)"; timeout 200 python tools/bench_nets.py --scans 16384 --chunk 1024 --steps 3 2>&1 | tail -2 | tee gpurun_out/nets_j.txt
echo "== RML_CG_RESIDENT=0"; RML_CG_RESIDENT=0 timeout 200 python tools/bench_nets.py --scans 16384 --chunk 1024 --steps 3 2>&1 | tail -2 | tee -a gpurun_out/nets_j.txt
date +%T
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k1_|k2_|k3|k4_|k5_" -c 40 --csv --log-file gpurun_out/launches_nets_r1j.csv python tools/bench_nets.py --scans 4096 --chunk 1024 --steps 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_nets_r1j.csv | cut -d'"' -f10,18,30 | sed -n '2,4p;22,27p'
date +%T
