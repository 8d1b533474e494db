#!/bin/bash
mkdir -p gpurun_out
echo "== racecheck (RML_T6_DBG=512: barrier form of the tower's mbarrier waits), without the bulk-copy projection kernels"
RML_T6_DBG=512 timeout 900 compute-sanitizer --tool racecheck --kernel-name-exclude kns=k1_project_max --error-exitcode 9 python tools/san_small.py > gpurun_out/sanitizer_racecheck_r3.log 2>&1
echo "rc=$?" >> gpurun_out/sanitizer_racecheck_r3.log
grep -v "Saved host\|Host Frame\|=========     at" gpurun_out/sanitizer_racecheck_r3.log | tail -8 | cut -c1-250
