# Round-1 GPU pass E: ncu full captures of the two heaviest network kernels (dnn, chunk 1024)
mkdir -p gpurun_out
date +%T
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k34_resize_conv1 -s 1 -c 1 -f -o gpurun_out/k34_full python tools/bench_nets.py --scans 2048 --chunk 1024 --steps 1 > /dev/null 2>&1; ls -la gpurun_out/k34_full.ncu-rep
date +%T
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k4_conv_igemm -s 1 -c 1 -f -o gpurun_out/k4_full python tools/bench_nets.py --scans 2048 --chunk 1024 --steps 1 > /dev/null 2>&1; ls -la gpurun_out/k4_full.ncu-rep
date +%T
timeout 200 python tools/bench_nets.py --scans 16384 --chunk 1024 --steps 3 2>&1 | tail -2
date +%T
