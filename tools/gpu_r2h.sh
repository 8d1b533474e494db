#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python tools/time_tower.py dnn; echo "--- split 49,50,49"; RML_T6_SPLIT=49,50,49 timeout 300 python tools/time_tower.py dnn; echo "--- sgan"; timeout 300 python tools/time_tower.py sgan_c) > gpurun_out/time_tower_r2h.txt 2>&1
grep -v Warn gpurun_out/time_tower_r2h.txt
