set -x
python -m pytest tests/test_gpu_config1.py -q -s 2>&1 | grep -E "config 1|passed|failed|FAILED|Error" | tail
H3D_PROFILE=step_device ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r02d.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_r02d.log 2>&1
python tools/summarise_launches.py gpurun_out/launches_r02d.csv | head -50
