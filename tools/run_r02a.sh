set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02a_pytest.log; tail -5 gpurun_out/r02a_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.log 2> gpurun_out/r02a_bench.err; tail -c 1500 gpurun_out/r02a_bench.log
for v in nll6 nll7; do
  H3D_LIB=$PWD/hic3defdr_b200/libh3d_$v.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02a_bench_$v.log 2>/dev/null
  grep -o '"ms_per_step": [0-9.]*\|"nll_ms": [0-9.]*\|"equalize_ms": [0-9.]*' gpurun_out/r02a_bench_$v.log
done
for a in 1 3 12; do
  H3D_QCML_AHEAD=$a python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02a_bench_ahead$a.log 2>/dev/null
  grep -o '"ms_per_step": [0-9.]*\|"nll_ms": [0-9.]*\|"equalize_ms": [0-9.]*' gpurun_out/r02a_bench_ahead$a.log
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02a_ref.log 2>/dev/null; tail -c 1200 gpurun_out/r02a_ref.log
nproc
