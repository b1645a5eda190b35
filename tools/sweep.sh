#!/bin/bash
# cfg-5 style sweep on one GPU: arenas x mode -> gpurun_out/sweep.txt (device value, e2e, k_roles, mlp)
mkdir -p gpurun_out; : > gpurun_out/sweep.txt
for team in 1 2 3; do for arenas in 1024 4096 16384 65536; do
  timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-ppo --team $team --arenas $arenas > gpurun_out/sw.json 2> gpurun_out/sw.err || { echo "team $team arenas $arenas FAILED: $(tail -2 gpurun_out/sw.err)" | tee -a gpurun_out/sweep.txt; continue; }
  python -c "
import json; b=json.load(open('gpurun_out/sw.json')); print('${team}v${team} arenas %6d  device %7.2fM  e2e %7.2fM  k_roles %.3f ms  mlp %.3f ms  obs %d' % ($arenas, b['value']/1e6, b['e2e']['value']/1e6, b['roofline']['launch_ms'], b['roofline_mlp']['launch_ms'], b['config']['obs_size']))" | tee -a gpurun_out/sweep.txt
done; done
