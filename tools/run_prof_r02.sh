set -x
H3D_PROFILE=step_device ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-id ::regex:"equalize_kernel|nll_kernel|lrt_fused|union_emit_staged|pool_pull|sort_pass|median_select|scale_filter":1 -o gpurun_out/prof_r02f -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_r02f.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
H3D_PROFILE=step_device ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r02f.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_r02f.log 2>&1
python tools/summarise_launches.py gpurun_out/launches_r02f.csv | head -30
