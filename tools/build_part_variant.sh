#!/bin/bash
# usage: tools/build_part_variant.sh NAME "-DNINT_PART_CTAS=4 ..." : relinks the library with the slab-list units (and
# nint_api.cu, which sizes the lists from the same constants) rebuilt with the extra flags; load it with NINT_LIB_VARIANT=NAME
set -e
cd "$(dirname "$0")/.."
name=$1; extra=$2
tmp=$(mktemp -d)
for u in nint_part_f64 nint_part_f32 nint_api; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
    -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Iinclude $extra -Xptxas -v -c ninterp_b200/csrc/$u.cu -o $tmp/$u.cu.o > $tmp/$u.log 2>&1 &
done
wait
grep -A2 "part_eval_kernelIdLb0ELi0ELb1ELb1\|part_scatter_kernelId\|part_merge_kernelId" $tmp/nint_part_f64.log | grep "Used\|spill" || true
objs=$(ls ninterp_b200/_build/*.o | grep -v "nint_part_\|nint_api")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ninterp_b200/libninterp_cuda_$name.so $objs $tmp/nint_part_f64.cu.o $tmp/nint_part_f32.cu.o $tmp/nint_api.cu.o
rm -rf $tmp
echo built ninterp_b200/libninterp_cuda_$name.so
