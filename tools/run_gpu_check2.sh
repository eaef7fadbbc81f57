set -x
python tools/microbench.py > gpurun_out/micro_new.log 2>&1; tail -1 gpurun_out/micro_new.log
H3D_LIB=$PWD/hic3defdr_b200/libh3d_old.so python tools/microbench.py > gpurun_out/micro_old.log 2>&1; tail -1 gpurun_out/micro_old.log
H3D_LIB=$PWD/hic3defdr_b200/libh3d_old.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*'
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*'
timeout 600 python bench.py --workload chr1_1kb --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_n1.log 2>gpurun_out/bench_cfg4_n1.err; tail -c 1800 gpurun_out/bench_cfg4_n1.log; tail -3 gpurun_out/bench_cfg4_n1.err
timeout 600 python bench.py --workload human5kb --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg3_n1.log 2>gpurun_out/bench_cfg3_n1.err; tail -c 1800 gpurun_out/bench_cfg3_n1.log; tail -3 gpurun_out/bench_cfg3_n1.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv; free -g | head -2
