#!/bin/bash
# r02g: the device PPO learner vs the torch restatement, the Learner loop, smoke, quick bench (both arms)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_learner.py tests/test_gpu_collector.py -m gpu -x -q -s > gpurun_out/pytest_r02g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02g.log
grep -v "^DiscreteAction\|^$" gpurun_out/pytest_r02g.log | tail -40 | cut -c1-600
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -4 gpurun_out/smoke.log | cut -c1-400
timeout 900 python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python -c "
import json; b=json.load(open('gpurun_out/bench_quick.json')); print('value %.3fM' % (b['value']/1e6), 'e2e %.3fM' % (b['e2e']['value']/1e6), 'e2e_collect %.3fM' % (b['e2e_collect']['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'mlp %.3f ms' % b['roofline_mlp']['launch_ms']); print(json.dumps(b.get('ppo_iteration'), indent=1))"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_quick.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python -c "
import json; b=json.load(open('gpurun_out/bench_ref_quick.json')); print('ref value %.3fM' % (b['value']/1e6), b['cpu_baseline']['sample']); print(b.get('ppo_iteration'))"
