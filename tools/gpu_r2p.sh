#!/bin/bash
# round 2 evidence: smoke, compute-sanitizer (memcheck, racecheck, synccheck), ncu full captures
mkdir -p gpurun_out
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool tools/san_small.py"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/san_small.py > gpurun_out/sanitizer_${tool}_r2.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_${tool}_r2.log
  tail -6 gpurun_out/sanitizer_${tool}_r2.log | cut -c1-200
done
echo "== ncu full: towers, igemm, dense stack"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k6_tower<64' -s 2 -c 1 -f -o gpurun_out/r2_k6_dnn python tools/bench_nets.py --scans 18944 --chunk 4736 --steps 1 > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none -k regex:k5_dense -s 1 -c 1 -f -o gpurun_out/r2_k5_dnn python tools/bench_nets.py --scans 18944 --chunk 4736 --steps 1 > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none -k regex:'k6_tower<128' -s 1 -c 1 -f -o gpurun_out/r2_k6_sgan python tools/bench_nets.py --scans 16384 --chunk 4096 --steps 1 > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none -k regex:k4_conv_igemm -s 2 -c 2 -f -o gpurun_out/r2_k4_sgan python tools/bench_nets.py --scans 16384 --chunk 4096 --steps 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none -k regex:k2_rbf_i8 -s 3 -c 1 -f -o gpurun_out/r2_k2 python bench.py --steps 1 --warmup 3 --skip-extras --scans-per-gpu 16384 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k[0-9]+_' -c 30 --csv --log-file gpurun_out/r2_launches_svc.csv python bench.py --steps 3 --warmup 3 --skip-extras --scans-per-gpu 65536 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
