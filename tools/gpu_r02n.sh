#!/bin/bash
# r02n: where the time of the penetration-depth search goes (sub-phase cycle counters), per-phase profile of the role kernel, no-EPA A/B
mkdir -p gpurun_out
RLG_B200_LIB=$PWD/build_ab/lib_epat.so timeout 300 python bench.py --steps 40 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/epat.err
grep "epa timing" gpurun_out/epat.err | tee gpurun_out/r02n_epa.txt
RLG_B200_LIB=$PWD/build_ab/lib_pt.so RLG_PHASE_DUMP=$PWD/gpurun_out/phase_prof.bin timeout 300 python bench.py --steps 40 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/pt.json 2> gpurun_out/pt.err
python tools/phase_prof.py gpurun_out/phase_prof.bin | tee gpurun_out/r02n_phase_prof.txt
for lib in build_ab/lib_noepa.so rlgymppo_cpp_b200/csrc/librlgym_b200.so; do
RLG_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 60 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('$lib', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'e2e %.3fM' % (b['e2e']['value']/1e6))" | tee -a gpurun_out/r02n_ab.txt
done
