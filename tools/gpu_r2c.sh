#!/bin/bash
# round 2, call C: ncu of the fused tower kernel (launch list + full capture with source)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[0-9]+_' -c 40 --csv --log-file gpurun_out/launches_nets_r2c.csv python tools/bench_nets.py --scans 4096 --chunk 2048 --steps 1 > /dev/null 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/launches_nets_r2c.csv') if not l.startswith('==')]
for row in csv.DictReader(lines):
    print(row['Kernel Name'][:70], row['Grid Size'], row['Block Size'], row['Metric Value'])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k6_tower -s 1 -c 1 -o gpurun_out/k6_full python tools/bench_nets.py --scans 4096 --chunk 2048 --steps 1 > /dev/null 2>&1
ls -la gpurun_out/k6_full.ncu-rep
