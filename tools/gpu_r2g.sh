#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python tools/prof_nets.py --scans 16384 --chunk 4096 --kind dnn; timeout 300 python tools/prof_nets.py --scans 32768 --chunk 8192 --kind dnn; timeout 300 python tools/prof_nets.py --scans 8192 --chunk 4096 --kind sgan_c) > gpurun_out/prof_nets_r2g.txt 2>&1
cat gpurun_out/prof_nets_r2g.txt | grep -v Warning
