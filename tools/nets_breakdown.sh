timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k1_|k3_|k4_|k5_" -c 80 --csv --log-file gpurun_out/launches_nets2.csv python tools/bench_nets.py --scans 8192 --chunk 1024 --steps 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_nets2.csv | python -c "
import csv,sys,collections
r=list(csv.reader(sys.stdin)); h=r[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
agg=collections.OrderedDict()
for row in r[1:]:
    k=row[ki][:30]+' '+row[gi]; agg.setdefault(k,[]).append(float(row[vi].replace(',','')))
for k,v in agg.items(): print(k, len(v), 'avg_us', round(sum(v)/len(v)/1e3,1), 'total_us', round(sum(v)/1e3,1))
"
