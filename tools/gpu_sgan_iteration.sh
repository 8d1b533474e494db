#!/bin/bash
# sgan iteration: parity of the network tests, then the per-kernel split of the default tree against the form it
# replaced, selected by $2 (an environment assignment, e.g. RML_K4_SHARE=0 or RML_T6_DBG=1024)
tag=${1:-r4a}
alt=${2:-RML_K4_SHARE=0}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py tests/test_gpu_fullsize.py -x -q > gpurun_out/pytest_nets_$tag.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_nets_$tag.log
tail -3 gpurun_out/pytest_nets_$tag.log | cut -c1-300
grep -n "^E  " gpurun_out/pytest_nets_$tag.log | head -5 | cut -c1-200
(echo "== default"; timeout 300 python tools/prof_nets.py --scans 8192 --chunk 4096 --kind sgan_c
 echo "== $alt"; env $alt timeout 300 python tools/prof_nets.py --scans 8192 --chunk 4096 --kind sgan_c) > gpurun_out/time_sgan_$tag.txt 2>&1
grep -v "Warn\|_warn" gpurun_out/time_sgan_$tag.txt
