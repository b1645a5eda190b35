#!/bin/bash
# r02x: inference overlapped with the tail of the fused step (programmatic dependent launch + per-block ready flags): tests + A/B
mkdir -p gpurun_out
tag=${1:-r02x}
timeout 1200 python -m pytest tests/test_gpu_collector.py tests/test_gpu_learner.py tests/test_gpu_ppo.py -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/pytest_$tag.log | head -20 | cut -c1-300
rm -f gpurun_out/${tag}_ab.txt
for i in 1 2; do for d in 0 1; do
RLG_COLLECT_OVERLAP=$d timeout 300 python bench.py --steps 100 --warmup 40 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); p=b.get('ppo_iteration') or {}; print('$tag overlap=$d', 'value %.3fM' % (b['value']/1e6), 'ms/step %.3f' % b['ms_per_step'], 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'mlp %.3f ms' % b['roofline_mlp']['launch_ms'], 'e2e %.3fM' % (b['e2e']['value']/1e6), 'ppo iter', p.get('total_iteration_time_s'), 'collect', p.get('collection_time_s'))" | tee -a gpurun_out/${tag}_ab.txt
done; done
