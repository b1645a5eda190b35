#!/bin/bash
mkdir -p gpurun_out
for v in "RLG_PDL=0" "RLG_COLLECT_OVERLAP=0" "RLG_PDL=1"; do
env $v CUDA_LAUNCH_BLOCKING=1 timeout 200 python bench.py --cfg4 --arenas 2048 --steps 3 --warmup 2 > gpurun_out/cfg4_dbg.json 2> gpurun_out/cfg4_dbg.err; echo "$v rc=$?"
grep -i "EngineError\|Error:" gpurun_out/cfg4_dbg.err | tail -2 | cut -c1-200
done
timeout 200 python - <<'PY' 2>&1 | tail -5
import numpy as np
from rlgymppo_cpp_b200 import abi, engine, collector
cfg = abi.default_cfg(num_arenas=2048, team_size=3)
e = engine.Engine(cfg); e.reset()
c = collector.Collector(e, max_steps=4, seed=3); c.init_default(seed=7)
try:
    c.collect(4); e.sync(); print("3v3 collect ok", c.read("action").shape)
except Exception as ex: print("3v3 collect FAILED", ex)
PY
