#!/bin/bash
# build_ab.sh <name> <nvcc -D flags...> — A/B build of the kernel objects with extra macros into ab_build/libgvdbx_<name>.so
# (same C ABI; select it with GVDBX_LIB=ab_build/libgvdbx_<name>.so).  The other objects come from the in-tree build.
set -e
name=$1; shift
cd "$(dirname "$0")/../gvdb-voxels_b200"
d=/tmp/ab_$name; rm -rf $d; mkdir -p $d ../ab_build
for f in gvdbx_api gvdbx_k_deep gvdbx_k_deepshadow gvdbx_k_voxel gvdbx_k_trilinear gvdbx_k_levelset gvdbx_k_extra; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --use_fast_math -Xcompiler -fPIC -Xptxas -v "$@" -c csrc/$f.cu -o $d/$f.o 2> $d/$f.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../ab_build/libgvdbx_$name.so $d/gvdbx_api.o build/gvdbx_multi.o build/gvdbx_host.o $d/gvdbx_k_extra.o \
     $d/gvdbx_k_levelset.o $d/gvdbx_k_trilinear.o $d/gvdbx_k_voxel.o $d/gvdbx_k_deep.o $d/gvdbx_k_deepshadow.o -lrt
grep -A2 "ILi3ELi0ELi64ELb1E" $d/gvdbx_k_deep.log | grep -i "regis" | sed "s/^/$name: /"
