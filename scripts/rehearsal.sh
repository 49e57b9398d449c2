mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench reference arm"; ( time timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/reh_ref.json 2> gpurun_out/reh_ref.err ) 2>&1 | grep real; tail -c 600 gpurun_out/reh_ref.json; echo
echo "== bench ours"; ( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/reh_ours.json 2> gpurun_out/reh_ours.err ) 2>&1 | grep real; tail -2 gpurun_out/reh_ours.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/reh_ours.json').read().strip().splitlines()[-1])
s=d.pop('secondary')
print({k:d[k] for k in ('value','ms_per_step','recall_at_100','gpu_launches')}, d['e2e'], d['roofline'], d['clocks'])
print({k:(v.get('queries_per_s') or v.get('ms') or v) for k,v in s.items() if isinstance(v,dict)})
print('c3_residual', s.get('c3_residual'))
PY
