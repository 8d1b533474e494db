#!/bin/bash
mkdir -p gpurun_out
for sp in 50,52,46 49,50,49 48,52,48 50,54,44 52,54,42 47,53,48 49,53,46 51,51,46; do
  echo -n "RML_T6_SPLIT=$sp  "; RML_T6_SPLIT=$sp timeout 200 python tools/time_tower.py dnn | tail -1
done > gpurun_out/tower_split_r3p.txt 2>&1
cat gpurun_out/tower_split_r3p.txt
