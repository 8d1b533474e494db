# Round-1 (third session) GPU pass: full -m gpu suite, bench, uint8-cube tests / sweep / ncu capture.
mkdir -p gpurun_out
date +%T
echo "== main gpu suite (without the uint8-cube file)"
timeout 800 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_u8cubes.py -p no:cacheprovider --durations=8 2>&1 | tail -16 | tee gpurun_out/pytest_gpu_main.log
date +%T
echo "== uint8-cube tests"
timeout 400 python -m pytest tests/test_gpu_u8cubes.py -q -p no:cacheprovider --timeout 150 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_u8.log
if grep -q " passed" gpurun_out/pytest_gpu_u8.log && ! grep -q "failed\|error\|Timeout" gpurun_out/pytest_gpu_u8.log; then U8OK=1; else U8OK=0; fi
echo "U8OK=$U8OK"; date +%T
echo "== bench"
if [ $U8OK = 1 ]; then
  timeout 500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"
else
  RML_BENCH_U8=0 timeout 500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"
fi
tail -c 2500 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
date +%T
if [ $U8OK = 1 ]; then
echo "== u8 sweep"
timeout 240 python tools/bench_u8.py 2>&1 | tail -12 | tee gpurun_out/u8_sweep.txt
echo "== ncu: launch list of the u8 pipeline + full capture of the u8 K1"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k1_|k2_" -c 12 --csv --log-file gpurun_out/launches_u8.csv python tools/bench_u8.py --scans 16384 --steps 2 --only-default > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_u8.csv | tail -5 | cut -c1-220
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_project_max_u8in -s 1 -c 1 -f -o gpurun_out/k1u8_full python tools/bench_u8.py --scans 16384 --steps 1 --only-default > /dev/null 2>&1; ls -la gpurun_out/k1u8_full.ncu-rep
fi
date +%T
echo "== nets: launch list at chunk 1024 (dnn)"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_nets_r1c.csv python tools/bench_nets.py --scans 4096 --chunk 1024 --steps 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_nets_r1c.csv | tail -40 | cut -d, -f5,9- | cut -c1-160
date +%T
