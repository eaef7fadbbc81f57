timeout 150 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
timeout 120 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_final_n2.log 2>gpurun_out/r02_final_n2.err
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_final_n2.log').read().strip().splitlines()[-1])
p=l['parity_vs_n1']
print('N=2', round(l['ms_per_step'],2), {k:round(v['ms'],2) for k,v in l['stages'].items()}, 'bit-identical', p['all_outputs_bit_identical'], p['disp_per_dist_max_rel'], p['n_q_lt_0_5'])
PY
grep -i "peer memory\|error" gpurun_out/r02_final_n2.err | head -3
