#!/bin/bash
# r02d: which of FMA contraction / approximate div+sqrt moves the GPU's contact ticks; single-TU speed of each
mkdir -p gpurun_out
for v in fma_precdiv nofma_approx inline_fast; do
RLG_B200_LIB=$PWD/build_ab/lib_$v.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "single_tick" > gpurun_out/pytest_parity_$v.log 2>&1
echo "== $v: $(grep -c '^E   ' gpurun_out/pytest_parity_$v.log) error lines; $(tail -1 gpurun_out/pytest_parity_$v.log)" | tee -a gpurun_out/r02d.txt
grep -n "^E       Assertion" gpurun_out/pytest_parity_$v.log | cut -c1-300 | tee -a gpurun_out/r02d.txt
RLG_B200_LIB=$PWD/build_ab/lib_$v.so timeout 300 python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('$v', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'])" | tee -a gpurun_out/r02d.txt
done
