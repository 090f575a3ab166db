#!/bin/bash
# Build a variant of the library with extra nvcc defines into tools/ab/<name>.so (A/B measurements via SAFEOPT_B200_LIB).
# usage: tools/build_variant.sh <name> [-DSO_K2_...=0 ...]
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/tools/ab; mkdir -p $out/obj_$name
for f in $(ls $root/safeopt_b200/csrc/*.cu | xargs -n1 basename | sed 's/\.cu$//'); do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
       -c $root/safeopt_b200/csrc/$f.cu -o $out/obj_$name/$f.o &
done
wait
nvcc -shared -o $out/$name.so $out/obj_$name/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
rm -rf $out/obj_$name
echo $out/$name.so
