#!/bin/bash
# round-2 profile evidence: full ncu capture of the dominant kernel, launch lists of the SVC step and the networks
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k1_project_max' -s 3 -c 1 -f -o gpurun_out/r2_k1_u8 python bench.py --steps 1 --warmup 3 --skip-extras --scans-per-gpu 16384 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[0-9]+_' -c 30 --csv --log-file gpurun_out/r2_launches_svc.csv python bench.py --steps 3 --warmup 3 --skip-extras --scans-per-gpu 65536 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[0-9]+_' -c 40 --csv --log-file gpurun_out/r2_launches_dnn.csv python tools/prof_nets.py --scans 16384 --chunk 8192 --kind dnn > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[0-9]+_' -c 40 --csv --log-file gpurun_out/r2_launches_sgan.csv python tools/prof_nets.py --scans 8192 --chunk 4096 --kind sgan_c > /dev/null 2>&1
ls -la gpurun_out/r2_k1_u8.ncu-rep gpurun_out/r2_launches_*.csv
