#!/bin/bash
# GPU gate + steady-state ncu capture of k_roles (+ launch list): tools/gpu_prof.sh [tag]
tag=${1:-cur}
mkdir -p gpurun_out
bash tools/gpu_quick.sh
timeout 800 ncu --set full --clock-control none --import-source on -k regex:k_roles -s 70 -c 1 -o gpurun_out/prof_k_roles_$tag -f python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-ppo > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ppo > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out | tail -8
