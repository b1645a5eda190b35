#!/bin/bash
# r02zp: global loads cached in L2 only (-Xptxas -dlcm=cg): does the L1 serve the per-thread stack better?
mkdir -p gpurun_out
rm -f gpurun_out/r02zp_ab.txt
for i in 1 2; do for lib in rlgymppo_cpp_b200/csrc/librlgym_b200.so build_ab/lib_cg.so; do
RLG_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 60 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('$lib', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'mlp %.3f' % b['roofline_mlp']['launch_ms'], 'e2e %.3fM' % (b['e2e']['value']/1e6))" | tee -a gpurun_out/r02zp_ab.txt
done; done
