#!/bin/bash
# sgan tower iteration: parity of the network tests, per-kernel split with the TMA-store epilogue and (RML_T6_DBG=1024) the
# warp-transpose + STG.128 form it replaced
tag=${1:-r4a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py -x -q > gpurun_out/pytest_nets_$tag.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_nets_$tag.log
tail -3 gpurun_out/pytest_nets_$tag.log | cut -c1-300
grep -n "^E  " gpurun_out/pytest_nets_$tag.log | head -5 | cut -c1-200
(echo "== TMA store epilogue"; timeout 300 python tools/prof_nets.py --scans 8192 --chunk 4096 --kind sgan_c
 echo "== RML_T6_DBG=1024 (warp transpose + STG.128)"; RML_T6_DBG=1024 timeout 300 python tools/prof_nets.py --scans 8192 --chunk 4096 --kind sgan_c) > gpurun_out/time_sgan_$tag.txt 2>&1
grep -v "Warn\|_warn" gpurun_out/time_sgan_$tag.txt
