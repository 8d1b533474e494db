#!/bin/bash
# diagnostic variants of k6_tower (wrong results, timing only): which part sets the period?
for d in 0 1 2 4 8 16 3 6 7 15 31; do
  echo -n "RML_T6_DBG=$d  "
  RML_T6_DBG=$d timeout 200 python tools/time_tower.py dnn 2>&1 | grep "n= 8192"
done
