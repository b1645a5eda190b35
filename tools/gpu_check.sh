#!/bin/bash
# Runs on the GPU box (via gpurun): GPU test-suite, smoke, bench, ncu launch list + one full capture of k_step.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 30 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json | cut -c1-600
# launch list (cold-cache, serialised): shares only
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ppo > gpurun_out/ncu_bench.log 2>&1
# one full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_roles -s 8 -c 1 -o gpurun_out/prof_k_roles python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ppo > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mlp_infer -s 12 -c 1 -o gpurun_out/prof_k_mlp python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ppo > gpurun_out/ncu_full_mlp.log 2>&1
ls -la gpurun_out
