timeout 150 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 240 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_final_n1.log 2>gpurun_out/r02_final_n1.err; tail -c 300 gpurun_out/r02_final_n1.err
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_final_n1.log').read().strip().splitlines()[-1])
print('N=1', round(l['ms_per_step'],2), 'value %.4g' % l['value'], l['e2e'], {k:round(v['ms'],2) for k,v in l['stages'].items()})
print('roofline', {k:(v if not isinstance(v,dict) else '...') for k,v in l['roofline'].items()})
print('hw', {k:v for k,v in (l['roofline'].get('hw') or {}).items() if not isinstance(v,dict)})
print('cpu', l['cpu_baseline']); print('clocks', l['clocks'], 'launches', l['gpu_launches'])
PY
H3D_REF_BINS=1200 timeout 120 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-700
