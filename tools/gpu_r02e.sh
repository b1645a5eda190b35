#!/bin/bash
mkdir -p gpurun_out
RLG_B200_LIB=$PWD/build_ab/lib_epat.so timeout 300 python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/epat.err
grep "epa timing" gpurun_out/epat.err | tee gpurun_out/r02e.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_parity.log 2>&1; tail -2 gpurun_out/pytest_parity.log
