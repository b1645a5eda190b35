#!/bin/bash
mkdir -p gpurun_out
bash tools/ab_bench.sh build_ab/lib_base.so build_ab/lib_fmad.so build_ab/lib_fast.so build_ab/lib_unroll.so 2>&1 | tee gpurun_out/ab1.txt
RLG_B200_LIB=$PWD/build_ab/lib_base.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_roles -s 70 -c 1 -o gpurun_out/prof_k_roles_steady python bench.py --steps 10 --warmup 10 --no-cpu-baseline --no-ppo > gpurun_out/ncu_full_steady.log 2>&1
tail -2 gpurun_out/ncu_full_steady.log | cut -c1-300
