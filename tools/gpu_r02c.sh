#!/bin/bash
# r02c: does IEEE arithmetic (no FMA contraction, exact div / sqrt) in the physics bring the GPU's contact ticks to the host build's parity?
mkdir -p gpurun_out
for lib in build_ab/lib_strict.so; do
RLG_B200_LIB=$PWD/$lib timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "single_tick" > gpurun_out/pytest_parity_strict.log 2>&1; echo "parity rc=$?" >> gpurun_out/pytest_parity_strict.log
grep -n "^E  \|passed\|failed" gpurun_out/pytest_parity_strict.log | cut -c1-400
cp gpurun_out/parity_r02_gpu.json gpurun_out/parity_r02_gpu_strict.json
done
