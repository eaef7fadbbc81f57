set -x
python -m pytest tests -m gpu -q -s 2>&1 | grep -E "config 1|equalize .*worst|passed|failed|FAILED|Error|end to end" | tail -30
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench.log 2>gpurun_out/r02c_bench.err; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02c_bench.log').read().strip().splitlines()[-1])
print(l['ms_per_step'], l['e2e'], {k:v['ms'] for k,v in l['stages'].items()}, l['qcml'])
PY
