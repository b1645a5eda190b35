#!/bin/bash
# r02zo: one-GPU lines of the other BASELINE workloads on the final code: configs[2] (2v2 padded + zero-sum, 8 192 arenas), 3v3 at 8 192 arenas
mkdir -p gpurun_out
rm -f gpurun_out/r02zo_ab.txt
run() { tag=$1; shift
timeout 300 python bench.py "$@" --steps 100 --warmup 40 --no-cpu-baseline > gpurun_out/r02zo_$tag.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/r02zo_$tag.json')); p=b.get('ppo_iteration') or {}; print('$tag', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'mlp %.3f ms' % b['roofline_mlp']['launch_ms'], 'e2e %.3fM' % (b['e2e']['value']/1e6), 'ppo iter', p.get('total_iteration_time_s'))" | tee -a gpurun_out/r02zo_ab.txt
}
run cfg3_1gpu --team 2 --padded-obs --zero-sum --arenas 8192
run 3v3_1gpu --team 3 --arenas 8192
run 2v2_1gpu --team 2 --arenas 16384
