#!/bin/bash
# r02zn: ncu --set full capture of k_roles (steady state) and k_mlp_infer on the final code
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:k_roles -s 70 -c 1 -o gpurun_out/prof_k_roles_r02zn -f python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-ppo > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mlp_infer -s 40 -c 1 -o gpurun_out/prof_k_mlp_r02zn -f python bench.py --steps 8 --warmup 5 --no-cpu-baseline --no-ppo > gpurun_out/ncu_full_mlp.log 2>&1
ls -la gpurun_out/*r02zn*
