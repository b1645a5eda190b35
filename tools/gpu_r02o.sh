#!/bin/bash
# r02o: penetration-depth search out of line (one copy of every helper): parity + timing + the search's cycle counters
mkdir -p gpurun_out
tag=${1:-r02o}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "single_tick or compiled_reference" > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/pytest_$tag.log | head -20 | cut -c1-300
RLG_B200_LIB=$PWD/build_ab/lib_epat.so timeout 300 python bench.py --steps 40 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/epat.err
grep "epa timing" gpurun_out/epat.err | tee gpurun_out/${tag}_epa.txt
for i in 1 2; do
timeout 300 python bench.py --steps 60 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
python -c "
import json; b=json.load(open('gpurun_out/ab.json')); print('$tag', 'value %.3fM' % (b['value']/1e6), 'k_roles %.3f ms' % b['roofline']['launch_ms'], 'e2e %.3fM' % (b['e2e']['value']/1e6))" | tee -a gpurun_out/${tag}_ab.txt
done
