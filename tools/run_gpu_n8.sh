set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n8.log 2>gpurun_out/bench_n8.err; tail -c 1400 gpurun_out/bench_n8.log
H3D_TRACE=1 timeout 300 $TR bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/trace_n8.log 2>&1; grep "h3d trace" gpurun_out/trace_n8.log | head -24
timeout 300 $TR bench.py --gpus 8 --shard rows --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_n8_rows.log 2>gpurun_out/bench_n8_rows.err; tail -c 1000 gpurun_out/bench_n8_rows.log
