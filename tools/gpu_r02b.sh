#!/bin/bash
# r02b: parity suite (batched tick fixtures -> parity_r02_gpu.json) + A/B of k_roles: round-1 library, current, current without the EPA search
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s > gpurun_out/pytest_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/pytest_parity.log
tail -4 gpurun_out/pytest_parity.log
for rep in 1 2; do
for lib in build_ab/lib_r01.so rlgymppo_cpp_b200/csrc/librlgym_b200.so build_ab/lib_noepa.so; do
  RLG_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('$lib', 'rep$rep', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'mlp %.3f ms' % b['roofline_mlp']['launch_ms'])" | tee -a gpurun_out/ab_r02b.txt
done; done
