# usage: run_ab.sh name1 name2 ...  (libh3d_<name>.so; "" = product build)
for v in "$@"; do
  lib=$PWD/hic3defdr_b200/libh3d${v:+_$v}.so
  [ "$v" = "main" ] && lib=$PWD/hic3defdr_b200/libh3d.so
  H3D_LIB=$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null > gpurun_out/ab_$v.log
  echo "$v: $(grep -o '"ms_per_step": [0-9.]*\|"nll_ms": [0-9.]*\|"equalize_ms": [0-9.]*' gpurun_out/ab_$v.log | tr '\n' ' ')"
done
