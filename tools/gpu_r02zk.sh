#!/bin/bash
# r02zk: balanced groups (4 x 28 lanes instead of 3 x 32 + 15 at 16 384 arenas): A/B
mkdir -p gpurun_out
rm -f gpurun_out/r02zk_ab.txt
for i in 1 2; do for d in 0 1; do
RLG_BALANCED_GROUPS=$d timeout 300 python bench.py --steps 60 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('balanced=$d', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'e2e %.3fM' % (b['e2e']['value']/1e6))" | tee -a gpurun_out/r02zk_ab.txt
done; done
RLG_BALANCED_GROUPS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "single_tick_random or sharded or many_arenas or full_size" 2>&1 | tail -2
