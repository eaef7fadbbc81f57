for v in "$@"; do
  lib=$PWD/hic3defdr_b200/libh3d${v:+_$v}.so
  [ "$v" = "main" ] && lib=$PWD/hic3defdr_b200/libh3d.so
  H3D_LIB=$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null > gpurun_out/ab_$v.log
  python - <<PY
import json
l=json.loads(open("gpurun_out/ab_$v.log").read().strip().splitlines()[-1])
print("$v", round(l["ms_per_step"],2), {k:round(v["ms"],2) for k,v in l["stages"].items() if k in ("prepare_data","lrt","bh","estimate_disp/qcml","estimate_disp/pool")})
PY
done
