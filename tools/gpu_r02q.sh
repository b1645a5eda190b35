#!/bin/bash
# r02q: penetration-depth searches deferred to the ball warps (EpaQueue): parity + A/B against the in-place evaluation (RLG_EPA_DEFER=0)
mkdir -p gpurun_out
tag=${1:-r02q}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/pytest_$tag.log | head -20 | cut -c1-300
for i in 1 2; do for d in 0 1; do
RLG_EPA_DEFER=$d timeout 300 python bench.py --steps 60 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('$tag defer=$d', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'e2e %.3fM' % (b['e2e']['value']/1e6))" | tee -a gpurun_out/${tag}_ab.txt
done; done
