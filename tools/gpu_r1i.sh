mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k34_resize_conv1 -s 1 -c 1 -f -o gpurun_out/k34_full_i python tools/bench_nets.py --scans 2048 --chunk 1024 --steps 1 > /dev/null 2>&1; ls -la gpurun_out/k34_full_i.ncu-rep
