#!/bin/bash
# r02zm: final numbers of round 2 on the final code (after the chained split step and the small-pool block shape): bench (+ cpu baseline, ppo iteration), reference arm, launch list
mkdir -p gpurun_out
tag=r02zm
timeout 900 python bench.py --steps 100 --warmup 40 > gpurun_out/bench_$tag.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ppo > gpurun_out/ncu_bench.log 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
python -c "
import json; b=json.load(open('gpurun_out/bench_$tag.json')); print('value %.3fM' % (b['value']/1e6), 'e2e %.3fM' % (b['e2e']['value']/1e6), 'e2e_collect %.3fM' % (b['e2e_collect']['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'mlp %.3f ms' % b['roofline_mlp']['launch_ms'], 'cpu', b['cpu_baseline']['value'], {k: v for k, v in b['ppo_iteration'].items() if k.endswith('_s')})
r=json.load(open('gpurun_out/bench_reference_$tag.json')); print('reference', r['value'], r.get('ppo_iteration'))"
