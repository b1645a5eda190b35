#!/bin/bash
# quick GPU gate: parity suite + smoke + short bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; b=json.load(open('gpurun_out/bench_quick.json')); print('value %.3fM' % (b['value']/1e6), 'e2e %.3fM' % (b['e2e']['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'mlp %.3f ms' % b['roofline_mlp']['launch_ms'])"
