#!/bin/bash
# r02zh: small pools (the strong-scaling floor): arenas per block below one full warp-group
mkdir -p gpurun_out
rm -f gpurun_out/r02zh_ab.txt
for cfg in "2048 0" "2048 14" "2048 16" "4096 0" "4096 28" "8192 0" "8192 56" "1024 0" "1024 7" "1024 8"; do
set -- $cfg
if [ $2 = 0 ]; then unset RLG_ARENAS_PER_BLOCK; else export RLG_ARENAS_PER_BLOCK=$2; fi
timeout 300 python bench.py --arenas $1 --steps 60 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('arenas $1 per block $2', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'mlp %.3f ms' % b['roofline_mlp']['launch_ms'], 'e2e %.3fM' % (b['e2e']['value']/1e6))" | tee -a gpurun_out/r02zh_ab.txt
done
