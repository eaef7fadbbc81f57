set -x
nvidia-smi -L | head -3
python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -15 > gpurun_out/r02_pytest_multi_n2.log; cat gpurun_out/r02_pytest_multi_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.log 2>gpurun_out/r02_bench_n2.err; tail -c 2500 gpurun_out/r02_bench_n2.log; tail -5 gpurun_out/r02_bench_n2.err | cut -c1-300
timeout 600 $TR bench.py --gpus 2 --shard rows --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_bench_n2_rows.log 2>gpurun_out/r02_bench_n2_rows.err; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_bench_n2_rows.log').read().strip().splitlines()[-1])
print('rows:', l['ms_per_step'], l['parity_vs_n1'])
PY
