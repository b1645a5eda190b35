#!/bin/bash
# r02j: P1 split (hitbox-mesh narrowphase on the ball warp): GPU parity + PPO tests + bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ppo.py tests/test_gpu_collector.py -m gpu -q > gpurun_out/pytest_r02j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02j.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/pytest_r02j.log | head -30 | cut -c1-300
cp profiles/parity_r02_gpu.json gpurun_out/parity_r02_gpu.json 2>/dev/null
timeout 900 python bench.py --steps 100 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/bench_quick.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python -c "
import json; b=json.load(open('gpurun_out/bench_quick.json')); print('value %.3fM' % (b['value']/1e6), 'median-of-blocks %.3fM' % (b['value_median_of_blocks']/1e6), 'e2e %.3fM' % (b['e2e']['value']/1e6), 'e2e_collect %.3fM' % (b['e2e_collect']['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'mlp %.3f ms' % b['roofline_mlp']['launch_ms'])"
timeout 300 python tools/ppo_prof.py --iters 8 2>&1 | grep -v Discrete | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ppo.csv python tools/ppo_prof.py --iters 2 > gpurun_out/ncu_ppo.log 2>&1
python tools/launch_shares.py gpurun_out/launches_ppo.csv 160 | head -24
