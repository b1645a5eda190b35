#!/bin/bash
# r02i: fixed PPO / shim tests, then the learning-curve A/B (engine 5 x 50 M, reference 5 x 20 M timesteps, run side by side)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo.py tests/test_cpp_shim.py -m gpu -q -s > gpurun_out/pytest_r02i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02i.log
grep -E "passed|failed|^FAILED|update cosine|rc=" gpurun_out/pytest_r02i.log | cut -c1-300
nproc
(timeout 1500 python tools/learning_curves.py --arm reference --seeds 5 --timesteps 20e6 --out gpurun_out/curves_reference.json > gpurun_out/curves_reference.log 2>&1; echo "ref rc=$?" >> gpurun_out/curves_reference.log) &
timeout 1500 python tools/learning_curves.py --arm engine --seeds 5 --timesteps 50e6 --out gpurun_out/curves_engine.json > gpurun_out/curves_engine.log 2>&1; echo "engine rc=$?" >> gpurun_out/curves_engine.log
wait
grep -v "^Learner\|Discrete" gpurun_out/curves_engine.log | tail -6 | cut -c1-400
grep -v "^Learner\|Discrete" gpurun_out/curves_reference.log | tail -6 | cut -c1-400
