set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python tools/microbench.py > gpurun_out/micro_new.log 2>&1; tail -1 gpurun_out/micro_new.log
python tools/time_class.py > gpurun_out/time_class.log 2>gpurun_out/time_class.err; tail -c 800 gpurun_out/time_class.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*'
