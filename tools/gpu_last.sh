#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_last.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_last.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/pytest_last.log | head -10 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 40 --warmup 40 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); p=b.get('ppo_iteration') or {}; print('last', 'value %.3fM' % (b['value']/1e6), 'e2e %.3fM' % (b['e2e']['value']/1e6), 'ppo iter %.3f ms' % (1e3*p.get('total_iteration_time_s',0)))"
