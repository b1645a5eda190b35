#!/bin/bash
# r02h: device PPO vs restatement (fixed tolerances), Learner loop, C++ shim: Learner app + host-plugin path bit-exactness
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_learner.py tests/test_cpp_shim.py -m gpu -q -s > gpurun_out/pytest_r02h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02h.log
grep -v "^DiscreteAction\|^$\|Lookup table" gpurun_out/pytest_r02h.log | grep -E "passed|failed|Error|error|assert|update cosine|mismatch|FATAL|rc=|^E " | head -60 | cut -c1-500
