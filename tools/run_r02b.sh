set -x
for v in "" _h2e4 _h3e5; do
  H3D_LIB=$PWD/hic3defdr_b200/libh3d$v.so python -m pytest tests/test_gpu_config1.py -q -s 2>&1 | grep -E "config 1|passed|failed|Error" 
  H3D_LIB=$PWD/hic3defdr_b200/libh3d$v.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep -o '"ms_per_step": [0-9.]*\|"nll_ms": [0-9.]*\|"equalize_ms": [0-9.]*'
done
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02b_pytest.log; tail -15 gpurun_out/r02b_pytest.log
