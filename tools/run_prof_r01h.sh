set -x
python bench.py > gpurun_out/bench_r01h.log 2>gpurun_out/bench_r01h.err; tail -c 300 gpurun_out/bench_r01h.log
H3D_PROFILE=step_device ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r01h.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
