set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_r01g.log 2>gpurun_out/bench_r01g.err; tail -c 600 gpurun_out/bench_r01g.log
H3D_PROFILE=step_device ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r01g.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
H3D_PROFILE=step_device ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-id ::regex:"equalize_kernel|nll_kernel|lrt_fused|union_emit_staged|gather_counts|bh_scatter|median_select|rank_emit":1 -o gpurun_out/prof_r01g -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
