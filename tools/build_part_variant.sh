#!/bin/bash
# usage: tools/build_part_variant.sh NAME "-DNINT_PART_CTAS=4 ..." : relinks the library with the slab-list units rebuilt
set -e
cd "$(dirname "$0")/.."
name=$1; extra=$2
tmp=$(mktemp -d)
for t in f64 f32; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
    -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Iinclude $extra -c ninterp_b200/csrc/nint_part_$t.cu -o $tmp/nint_part_$t.cu.o &
done
wait
objs=$(ls ninterp_b200/_build/*.o | grep -v nint_part_)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ninterp_b200/libninterp_cuda_$name.so $objs $tmp/nint_part_f64.cu.o $tmp/nint_part_f32.cu.o
rm -rf $tmp
echo built ninterp_b200/libninterp_cuda_$name.so
