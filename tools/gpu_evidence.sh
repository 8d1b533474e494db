# Round evidence: smoke(), compute-sanitizer on a small pass, ncu launch list + full captures of the
# hot kernels on an 8192-scan batch (fast under replay), nets at the configs[2]/[4] sizes.
mkdir -p gpurun_out
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== sanitizer (memcheck) on the pipeline + nets tests"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "predict_pipeline or k1_max_u8 or golden_reference" 2>&1 | tail -4
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k1_|k2_" -c 24 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 3 --warmup 3 --skip-extras --scans-per-gpu 16384 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_r1b.csv | tail -6 | cut -c1-200
echo "== ncu full K1 / K2"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_project_max -s 3 -c 1 -f -o gpurun_out/k1_full_r1b python bench.py --steps 1 --warmup 3 --skip-extras --scans-per-gpu 16384 > /dev/null 2>&1; ls -la gpurun_out/k1_full_r1b.ncu-rep
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k2_rbf_i8 -s 3 -c 1 -f -o gpurun_out/k2_full_r1b python bench.py --steps 1 --warmup 3 --skip-extras --scans-per-gpu 16384 > /dev/null 2>&1; ls -la gpurun_out/k2_full_r1b.ncu-rep
echo "== nets at config sizes"
timeout 300 python tools/bench_nets.py --scans 262144 --chunk 1024 --steps 2 2>&1 | tail -2
