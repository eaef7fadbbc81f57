for ns in 2 4 8 12; do
H3D_PREPARE_STREAMS=$ns python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams $ns', l['ms_per_step'], l['stages']['prepare_data']['ms'])"
done
