#!/bin/bash
# A/B helper: builds hic3defdr_b200/libh3d_<name>.so with extra nvcc flags
# (select it at run time with H3D_LIB=<path>).  Not part of the product build.
set -e
name=$1; shift
cd "$(dirname "$0")/.."
out=hic3defdr_b200/build_$name
mkdir -p $out
for f in hic3defdr_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $f -o $out/$b.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o hic3defdr_b200/libh3d_$name.so $out/*.o
echo built hic3defdr_b200/libh3d_$name.so
