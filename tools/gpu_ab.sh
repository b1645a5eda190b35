#!/bin/bash
# bench-only A/B on one box: tools/gpu_ab.sh "ENV1=.. ENV2=.." "ENV=.." ...   (each arg = env assignments for one variant; "-" = none)
# extra bench flags via BENCH_FLAGS
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  envs=""; [ "$v" != "-" ] && envs="$v"
  env $envs timeout 300 python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-ppo $BENCH_FLAGS > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('[$v]', 'rep$rep', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'mlp %.3f ms' % b['roofline_mlp']['launch_ms'])"
done; done
