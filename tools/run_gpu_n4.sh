set -x
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
timeout 300 $TR4 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.log 2>gpurun_out/bench_n4.err; grep -o '"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*' gpurun_out/bench_n4.log
timeout 300 $TR2 bench.py --gpus 2 --shard rows --steps 5 --warmup 3 > gpurun_out/bench_n2_rows.log 2>gpurun_out/bench_n2_rows.err; grep -o '"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*' gpurun_out/bench_n2_rows.log; tail -3 gpurun_out/bench_n2_rows.err | cut -c1-300
timeout 300 $TR4 bench.py --impl reference --gpus 4 --steps 1 --warmup 1 2>/dev/null | cut -c1-300
