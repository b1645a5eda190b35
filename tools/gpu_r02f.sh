#!/bin/bash
# r02f: the tightened rollout-statistics test + the device state-setter distributions on the GPU, then a quick bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "rollout or state_setters" > gpurun_out/pytest_r02f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02f.log
tail -30 gpurun_out/pytest_r02f.log | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; b=json.load(open('gpurun_out/bench_quick.json')); print('value %.3fM' % (b['value']/1e6), 'e2e %.3fM' % (b['e2e']['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'mlp %.3f ms' % b['roofline_mlp']['launch_ms']); print(b.get('ppo_iteration'))"
