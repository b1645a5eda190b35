#!/bin/bash
# r02zl: the split host-buffer step's second launch as the programmatic dependent of the first (per-block flags): parity + e2e A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "step_host or sharded or gym_layer or reward_metrics" 2>&1 | tail -2
rm -f gpurun_out/r02zl_ab.txt
for i in 1 2; do for d in 0 1; do
RLG_SPLIT_CHAIN=$d timeout 300 python bench.py --steps 40 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('split chain=$d', 'value %.3fM' % (b['value']/1e6), 'e2e %.3fM' % (b['e2e']['value']/1e6), 'checksum', b['e2e']['reward_checksum'])" | tee -a gpurun_out/r02zl_ab.txt
done; done
