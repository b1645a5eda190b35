#!/bin/bash
# A/B the k_step kernel across library variants in ONE box (same clocks): tools/ab_bench.sh build_ab/lib_A.so build_ab/lib_B.so ...
mkdir -p gpurun_out
for rep in 1 2; do
for lib in "$@"; do
  RLG_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('$lib', 'rep$rep', 'value %.3fM' % (b['value']/1e6), 'k_step %.3f ms' % b['roofline']['launch_ms'], 'mlp %.3f ms' % b['roofline_mlp']['launch_ms'])"
done; done
