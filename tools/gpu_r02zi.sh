#!/bin/bash
# r02zi: power-of-two sub-warp groups for small pools: full GPU suite + small-pool bench lines
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_r02zi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02zi.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/pytest_r02zi.log | head -20 | cut -c1-300
rm -f gpurun_out/r02zi_ab.txt
for a in 1024 2048 4096 16384; do
timeout 300 python bench.py --arenas $a --steps 60 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('arenas $a', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'e2e %.3fM' % (b['e2e']['value']/1e6))" | tee -a gpurun_out/r02zi_ab.txt
done
