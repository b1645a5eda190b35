#!/bin/bash
# r02zc: user ActionParser tables (rlg_engine_set_action_table), PDL fallback path, full GPU suite
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_r02zc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02zc.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/pytest_r02zc.log | head -20 | cut -c1-300
timeout 300 python bench.py --steps 60 --warmup 40 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); p=b.get('ppo_iteration') or {}; print('r02zc', 'value %.3fM' % (b['value']/1e6), 'ms/step %.3f' % b['ms_per_step'], 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'e2e %.3fM' % (b['e2e']['value']/1e6), 'ppo iter', p.get('total_iteration_time_s'))"
