#!/bin/bash
# Developer tool: build the library with extra -D switches as _var/libwsr_<name>.so (see ab_variants.sh)
# usage: tools/build_variant.sh <name> [-DFOO=1 ...]
set -e
name=$1; shift
cd "$(dirname "$0")/../wiser_b200/csrc"
mkdir -p ../../_var /tmp/var_$name
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 -Xcompiler -pthread -I../../include"
for f in wsr_capi kernels frontend; do nvcc $FLAGS "$@" -c -o /tmp/var_$name/$f.o $f.cu & done
nvcc $FLAGS "$@" -x cu -c -o /tmp/var_$name/host_index.o host_index.cc &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../_var/libwsr_$name.so /tmp/var_$name/*.o -lpthread -ldl
ls -la ../../_var/libwsr_$name.so
