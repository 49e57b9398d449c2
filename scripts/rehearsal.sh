mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench ours"; ( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/reh_ours.json 2> gpurun_out/reh_ours.err ) 2>&1 | grep real; tail -1 gpurun_out/reh_ours.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/reh_ours.json').read().strip().splitlines()[-1])
s=d.pop('secondary')
print({k:d[k] for k in ('value','ms_per_step','recall_at_100','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['scan_share_of_step'], d['clocks'])
print({k:(v.get('queries_per_s') or v.get('ms') or str(v)[:80]) for k,v in s.items() if isinstance(v,dict)})
PY
