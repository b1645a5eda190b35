#!/bin/bash
# r02zg: the PPO update's launches as programmatic dependents (csrc/pdl.h): tests + A/B of the PPO iteration (RLG_PDL=0/1)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_gemm.py tests/test_gpu_learner.py -m gpu -q -x > gpurun_out/pytest_r02zg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02zg.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/pytest_r02zg.log | head -10 | cut -c1-300
rm -f gpurun_out/r02zg_ab.txt
for i in 1 2; do for d in 0 1; do
RLG_PDL=$d timeout 300 python bench.py --steps 40 --warmup 40 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); p=b.get('ppo_iteration') or {}; print('r02zg RLG_PDL=$d', 'value %.3fM' % (b['value']/1e6), 'ppo iter %.3f ms' % (1e3*p.get('total_iteration_time_s',0)), 'learn %.3f ms' % (1e3*p.get('ppo_learn_time_s',0)), 'learn device %.3f ms' % (1e3*p.get('ppo_learn_device_time_s',0)), 'cdl iter %.3f ms' % (1e3*(p.get('collection_during_learn') or {}).get('total_iteration_time_s',0)))" | tee -a gpurun_out/r02zg_ab.txt
done; done
RLG_PDL=0 timeout 300 python tools/gemm_bench.py > gpurun_out/r02zg_gemm_pdl0.txt 2>&1; timeout 300 python tools/gemm_bench.py > gpurun_out/r02zg_gemm_pdl1.txt 2>&1; tail -4 gpurun_out/r02zg_gemm_pdl0.txt; tail -4 gpurun_out/r02zg_gemm_pdl1.txt
