set -x
python tools/microbench.py 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --workload chr1_1kb --shard rows --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_n2_rows.log 2>gpurun_out/bench_cfg4_n2_rows.err; tail -c 1500 gpurun_out/bench_cfg4_n2_rows.log; tail -5 gpurun_out/bench_cfg4_n2_rows.err
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.log 2>gpurun_out/bench_n2.err; tail -c 1500 gpurun_out/bench_n2.log
timeout 600 $TR bench.py --gpus 2 --shard rows --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_rows.log 2>gpurun_out/bench_n2_rows.err; tail -c 1500 gpurun_out/bench_n2_rows.log; tail -5 gpurun_out/bench_n2_rows.err
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
