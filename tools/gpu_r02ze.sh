#!/bin/bash
# r02ze: hitbox narrowphase reads the car's own work state instead of the global hand-over record: parity + timing (two bench runs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_r02ze.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02ze.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/pytest_r02ze.log | head -10 | cut -c1-300
rm -f gpurun_out/r02ze_ab.txt
for i in 1 2 3; do
timeout 300 python bench.py --steps 100 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('r02ze', 'value %.3fM' % (b['value']/1e6), 'ms/step %.3f' % b['ms_per_step'], 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'e2e %.3fM' % (b['e2e']['value']/1e6))" | tee -a gpurun_out/r02ze_ab.txt
done
