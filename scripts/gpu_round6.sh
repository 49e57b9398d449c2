mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -4
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"coarse_gemm|probe_select|ivfpq_scan|merge_topk|col_sqnorm|normalize_columns|lut_scan" -c 120 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-secondary --cpu-sample 0.2 > gpurun_out/r02_launches_c3.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_c3.csv')) if len(r)>5]
h=[i for i,r in enumerate(rows) if r[0]=='ID'][0]; H=rows[h]; D=rows[h+1:]
ki=H.index('Kernel Name'); vi=H.index('Metric Value')
seq=[(r[ki].split('(')[0][-40:], float(r[vi].replace(',',''))/1e6) for r in D]
for n,v in seq[-14:]: print('%-42s %8.3f ms'%(n,v))
PY
