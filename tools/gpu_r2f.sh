#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py -x -q > gpurun_out/pytest_nets_r2f.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_nets_r2f.log
tail -6 gpurun_out/pytest_nets_r2f.log | cut -c1-300
grep -n "^E  " gpurun_out/pytest_nets_r2f.log | head -5 | cut -c1-200
timeout 300 python tools/bench_nets.py --scans 16384 --chunk 4096 > gpurun_out/nets_r2f.txt 2>&1
cat gpurun_out/nets_r2f.txt | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[0-9]+_' -c 24 --csv --log-file gpurun_out/launches_nets_r2f.csv python tools/bench_nets.py --scans 4096 --chunk 2048 --steps 1 > /dev/null 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/launches_nets_r2f.csv') if not l.startswith('==')]
for row in csv.DictReader(lines):
    print(row['Kernel Name'][:70], row['Grid Size'], row['Block Size'], row['Metric Value'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k6_tower -s 1 -c 1 -o gpurun_out/k6_full_e python tools/bench_nets.py --scans 4096 --chunk 2048 --steps 1 > /dev/null 2>&1
ls -la gpurun_out/k6_full_e.ncu-rep
