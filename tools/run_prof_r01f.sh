set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_r01f.log 2>gpurun_out/bench_r01f.err; tail -c 600 gpurun_out/bench_r01f.log
H3D_PROFILE=step_device ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
H3D_PROFILE=step_device ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"equalize|nll_kernel|lrt_fused" -c 3 -o gpurun_out/prof_r01f -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
