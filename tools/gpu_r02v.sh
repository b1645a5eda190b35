#!/bin/bash
# r02v: two half-size blocks per SM (a search stalls 55 arenas instead of 111) vs one block per SM; 65 536 arenas on one GPU (strong-scaling table)
mkdir -p gpurun_out
rm -f gpurun_out/r02v_ab.txt
for apb in 0 56 74; do
if [ $apb = 0 ]; then unset RLG_ARENAS_PER_BLOCK; else export RLG_ARENAS_PER_BLOCK=$apb; fi
timeout 300 python bench.py --steps 60 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('arenas per block $apb', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'e2e %.3fM' % (b['e2e']['value']/1e6))" | tee -a gpurun_out/r02v_ab.txt
done
unset RLG_ARENAS_PER_BLOCK
timeout 300 python bench.py --total-arenas 65536 --steps 60 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/r02v_strong64k_1.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/r02v_strong64k_1.json')); print('65536 arenas on one GPU', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'e2e %.3fM' % (b['e2e']['value']/1e6))" | tee -a gpurun_out/r02v_ab.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_cpp_shim.py -m gpu -q -x -k "state_setter or custom or host_matches or sharded or plugins" 2>&1 | tail -3
