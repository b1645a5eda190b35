#!/bin/bash
# tools/build_variant.sh NAME [extra nvcc flags...] -> build_ab/lib_NAME.so (A/B variants of the library; see tools/ab_bench.sh)
set -e
name=$1; shift
mkdir -p build_ab
FM="-fmad=false"
for f in "$@"; do case "$f" in -fmad=*) FM="";; esac; done
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 $FM --expt-relaxed-constexpr -Xcompiler -fPIC -shared \
  -Iinclude "$@" rlgymppo_cpp_b200/csrc/*.cu -o build_ab/lib_$name.so
echo built build_ab/lib_$name.so
