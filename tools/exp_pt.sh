#!/bin/bash
# quick gate + per-phase cycle profile (diagnostic build build_ab/lib_pt.so)
mkdir -p gpurun_out
bash tools/gpu_quick.sh 2>&1 | tail -8
RLG_B200_LIB=$PWD/build_ab/lib_pt.so RLG_PHASE_DUMP=$PWD/gpurun_out/phase_prof.bin timeout 300 python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-ppo > gpurun_out/pt.json 2> gpurun_out/pt.err
python tools/phase_prof.py gpurun_out/phase_prof.bin | tee gpurun_out/phase_prof.txt
