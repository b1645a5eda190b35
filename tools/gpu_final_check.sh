#!/bin/bash
# last check of the round: full GPU suite, then tools/gpu_r02zm.sh (bench, reference arm, launch list, smoke)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/pytest_final.log | head -10 | cut -c1-300
bash tools/gpu_r02zm.sh
