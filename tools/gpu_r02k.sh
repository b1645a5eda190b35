#!/bin/bash
# r02k: A/B of the hitbox offload (same box), GEMM epilogue rewrite: gemm tests + ppo profile
mkdir -p gpurun_out
for off in 0 2 0 2; do
RLG_HB_OFFLOAD=$off timeout 300 python bench.py --steps 60 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('offload $off', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'])" | tee -a gpurun_out/r02k_ab.txt
done
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_ppo.py -m gpu -q > gpurun_out/pytest_r02k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02k.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/pytest_r02k.log | head -20 | cut -c1-300
timeout 300 python tools/ppo_prof.py --iters 8 2>&1 | grep -v Discrete | tail -2
timeout 300 python tools/gemm_bench.py 2>&1 | tail -6
