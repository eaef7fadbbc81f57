set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.log 2>gpurun_out/r02_bench_n8.err; tail -5 gpurun_out/r02_bench_n8.err | cut -c1-300
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_n8.log').read().strip().splitlines()[-1])
print('N=8', l['ms_per_step'], l['value'], l['e2e'], {k:v['ms'] for k,v in l['stages'].items()}, l['qcml'], l['parity_vs_n1'])
PY
