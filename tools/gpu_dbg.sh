#!/bin/bash
# which role bounds k6_tower<64>: time the towers with parts of the work switched off (results wrong)
mkdir -p gpurun_out
for d in 0 2 29 31 1 4 8; do
  echo -n "RML_T6_DBG=$d  "; RML_T6_DBG=$d timeout 200 python tools/time_tower.py dnn | tail -1
done > gpurun_out/tower_dbg_${1:-x}.txt 2>&1
cat gpurun_out/tower_dbg_${1:-x}.txt
