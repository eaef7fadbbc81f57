set -x
python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -15 > gpurun_out/r02_pytest_multi_n2.log; tail -4 gpurun_out/r02_pytest_multi_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
for mode in peer nccl; do
H3D_EXCHANGE=$mode timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/r02_bench_n2_$mode.log 2>gpurun_out/r02_bench_n2_$mode.err; tail -3 gpurun_out/r02_bench_n2_$mode.err | cut -c1-300
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_n2_$mode.log').read().strip().splitlines()[-1])
print('$mode', l['ms_per_step'], {k:v['ms'] for k,v in l['stages'].items()}, l['parity_vs_n1']['all_outputs_bit_identical'])
PY
done

