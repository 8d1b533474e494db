# Round-end style GPU pass: the whole -m gpu suite, smoke(), the N=1 bench (what the driver runs).
mkdir -p gpurun_out
date +%T
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_all.log
date +%T
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
date +%T
timeout 400 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n1_final.json'))
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'clocks', d['clocks'])
print('parity', d['parity']); print('u8', d['u8_cubes']['value'], d['u8_cubes']['e2e']['value'], d['u8_cubes']['labels_equal_f32_path'])"
date +%T
