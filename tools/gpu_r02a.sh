#!/bin/bash
# r02a: GPU suite with the EPA narrowphase + tightened gates, smoke, bench (regression check of k_roles)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s > gpurun_out/pytest_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/pytest_parity.log
tail -5 gpurun_out/pytest_parity.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 30 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json | cut -c1-1800
ls -la gpurun_out
