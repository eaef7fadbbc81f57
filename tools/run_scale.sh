# usage: run_scale.sh N
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.log 2>gpurun_out/r02_bench_n$N.err; tail -2 gpurun_out/r02_bench_n$N.err | cut -c1-300
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_n$N.log').read().strip().splitlines()[-1])
p=l['parity_vs_n1']
print('N=$N', round(l['ms_per_step'],2), 'value %.3g' % l['value'], 'e2e %.3g' % l['e2e']['value'], {k:round(v['ms'],2) for k,v in l['stages'].items()}, {k:round(v,1) if isinstance(v,float) else v for k,v in l['qcml'].items()}, 'bit-identical', p['all_outputs_bit_identical'], p['disp_per_dist_max_rel'])
PY
