mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -8
timeout 600 python scripts/bench_aux.py centroids 2>&1 | grep -v "^$" | tail -1
timeout 900 python bench.py --steps 10 > gpurun_out/bench_r4.json 2> gpurun_out/bench_r4.err; tail -2 gpurun_out/bench_r4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r4.json').read().strip().splitlines()[-1])
print('c3 qps %.0f e2e %.0f scan_frac %.3f'%(d['value'],d['e2e']['value'],d['roofline']['frac']))
s=d['secondary']
for k in ('c2','c4','c3c'):
    print(k, 'qps %.0f frac %.3f step_frac %.3f recall %.3f'%(s[k]['queries_per_s'],s[k]['roofline']['frac'],s[k]['roofline']['step_frac_of_peak'],s[k]['recall_at_100']))
print(s['c5_assign'], s['c5_compute_centroids'], s['add_1M'])
PY
