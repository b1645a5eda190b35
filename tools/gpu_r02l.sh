#!/bin/bash
# r02l: non-blocking EPA workspaces: parity + timing (A/B against the r02k build is in profiles/r02k_ab.txt: 1.307 ms)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "single_tick or compiled_reference" > gpurun_out/pytest_r02l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02l.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/pytest_r02l.log | head -20 | cut -c1-300
for i in 1 2; do
timeout 300 python bench.py --steps 60 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('epa-nonblocking', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'e2e %.3fM' % (b['e2e']['value']/1e6))" | tee -a gpurun_out/r02l_ab.txt
done
