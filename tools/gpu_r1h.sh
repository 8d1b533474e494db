# Round-1 GPU pass H: resize passes with register-resident coefficients
mkdir -p gpurun_out
date +%T
timeout 400 python -m pytest tests/test_gpu_nets.py tests/test_gpu_u8cubes.py -q -p no:cacheprovider -k "nets or net_" --timeout 200 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_nets.log
date +%T
timeout 200 python tools/bench_nets.py --scans 16384 --chunk 1024 --steps 3 2>&1 | tail -2 | tee gpurun_out/nets_h.txt
date +%T
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k1_|k2_|k3|k4_|k5_" -c 40 --csv --log-file gpurun_out/launches_nets_r1h.csv python tools/bench_nets.py --scans 4096 --chunk 1024 --steps 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_nets_r1h.csv | cut -d'"' -f10,18,30 | sed -n '2,11p;22,27p'
date +%T
