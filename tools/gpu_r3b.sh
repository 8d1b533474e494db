#!/bin/bash
mkdir -p gpurun_out
for d in 0 128 256; do echo -n "RML_T6_DBG=$d (poll back-off 40 / 0 / 200 ns)  "; RML_T6_DBG=$d timeout 200 python tools/time_tower.py dnn | tail -1; done > gpurun_out/tower_poll_r3b.txt 2>&1
cat gpurun_out/tower_poll_r3b.txt
