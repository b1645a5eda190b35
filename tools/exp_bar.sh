#!/bin/bash
# barrier-scope / async-load A/B + per-phase cycle profile (diagnostic build build_ab/lib_pt.so)
mkdir -p gpurun_out
bash tools/gpu_ab.sh - RLG_BARRIER_MODE=1 RLG_BARRIER_MODE=2 RLG_ASYNC_LOAD=1 "RLG_BARRIER_MODE=1 RLG_ASYNC_LOAD=1" 2>&1 | tee gpurun_out/exp_bar.txt
for m in 0 1; do
RLG_BARRIER_MODE=$m RLG_B200_LIB=$PWD/build_ab/lib_pt.so RLG_PHASE_DUMP=$PWD/gpurun_out/phase_prof_b$m.bin timeout 300 python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-ppo > gpurun_out/pt$m.json 2> gpurun_out/pt$m.err
python tools/phase_prof.py gpurun_out/phase_prof_b$m.bin | tee gpurun_out/phase_prof_b$m.txt
done
