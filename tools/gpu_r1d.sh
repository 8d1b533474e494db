# Round-1 GPU pass D: uint8 K1 with two CTAs per SM + paired slabs; nets launch list; bench.
mkdir -p gpurun_out
date +%T
echo "== uint8-cube tests"
timeout 400 python -m pytest tests/test_gpu_u8cubes.py -q -p no:cacheprovider --timeout 150 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_u8.log
date +%T
echo "== u8 sweep (2 CTAs/SM)"
timeout 240 python tools/bench_u8.py 2>&1 | tail -8 | tee gpurun_out/u8_sweep_d.txt
echo "== u8 sweep (1 CTA/SM)"
RML_K1U8_CTAS=1 timeout 240 python tools/bench_u8.py --splits 24 2>&1 | tail -4 | tee gpurun_out/u8_sweep_d_1cta.txt
date +%T
echo "== ncu full capture of the u8 K1"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_project_max_u8in -s 1 -c 1 -f -o gpurun_out/k1u8_full_d python tools/bench_u8.py --scans 16384 --steps 1 --only-default > /dev/null 2>&1; ls -la gpurun_out/k1u8_full_d.ncu-rep
date +%T
echo "== nets: launch list at chunk 1024"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k1_|k2_|k3|k4_|k5_" -c 80 --csv --log-file gpurun_out/launches_nets_r1d.csv python tools/bench_nets.py --scans 4096 --chunk 1024 --steps 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_nets_r1d.csv | cut -d'"' -f10,14,16,30 | tail -50
date +%T
echo "== bench"
timeout 500 python bench.py > gpurun_out/bench_n1_d.json 2> gpurun_out/bench_n1_d.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n1_d.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
print('u8', json.dumps(d.get('u8_cubes'))[:900])"
date +%T
