#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "second_smaller or sharded" 2>&1 | tail -2
timeout 300 python bench.py --cfg4 --arenas 8192 --steps 12 --warmup 4 > gpurun_out/r02zr_cfg4_1gpu.json 2> gpurun_out/cfg4.err; echo "cfg4 rc=$?"
python -c "
import json; b=json.loads(open('gpurun_out/r02zr_cfg4_1gpu.json').read().strip().splitlines()[-1]); print('cfg4 1 GPU: value %.2fM' % (b['value']/1e6), 'ms/iteration %.3f' % b['ms_per_step'], b['iteration'])"
