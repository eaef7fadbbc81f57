set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.log 2>gpurun_out/bench_n8.err; grep -o '"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*\|"value": [0-9.]*' gpurun_out/bench_n8.log | head -3; tail -2 gpurun_out/bench_n8.err | cut -c1-200
