set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g.log 2>gpurun_out/bench_g.err; tail -c 1500 gpurun_out/bench_g.log
H3D_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/trace_g.log 2>&1
python tools/time_class.py > gpurun_out/time_class.log 2>gpurun_out/time_class.err; tail -c 800 gpurun_out/time_class.log
