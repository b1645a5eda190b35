#!/bin/bash
# r02z: final round-2 evidence: full GPU suite, smoke, then tools/gpu_final.sh (bench, reference arm, launch list, ncu captures)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r02z.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02z.log
tail -5 gpurun_out/pytest_r02z.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
bash tools/gpu_final.sh r02z
