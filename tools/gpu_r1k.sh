mkdir -p gpurun_out
for m in 3 4; do
echo "== RML_CG_DEBUG=$m"
RML_CG_DEBUG=$m timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k4_conv_igemm" -c 6 --csv --log-file gpurun_out/cg_debug_$m.csv python tools/bench_nets.py --scans 2048 --chunk 1024 --steps 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/cg_debug_$m.csv | cut -d'"' -f10,18,30 | tail -6
done
