#!/bin/bash
# soak: long runs of the overlapped paths (programmatic dependent launches + flag waits) under a timeout: collect loop, host-buffer step, learner
mkdir -p gpurun_out
timeout 600 python bench.py --steps 4000 --warmup 40 --no-cpu-baseline --no-ppo > gpurun_out/soak_bench.json 2> gpurun_out/soak.err; echo "long bench rc=$?"
python -c "
import json; b=json.load(open('gpurun_out/soak_bench.json')); print('4000 collects: value %.3fM' % (b['value']/1e6), 'e2e %.3fM' % (b['e2e']['value']/1e6), 'blocks', [round(x/1e6,2) for x in b['block_rates']])"
timeout 600 python - <<'PY'
import time, numpy as np
from rlgymppo_cpp_b200 import abi, learner as L
cfg = L.LearnerConfig(timestepsPerIteration=4096*2*4, expBufferSize=4096*2*4*2, randomSeed=3, sendMetrics=False, checkpointLoadFolder="", checkpointSaveFolder="", collectionDuringLearn=True)
cfg.ppo = L.PPOLearnerConfig(batchSize=4096*2*4, miniBatchSize=4096*2, epochs=2)
lr = L.Learner(abi.default_cfg(num_arenas=4096, team_size=1), cfg)
t0 = time.time(); reps = lr.learn(max_iterations=400); dt = time.time() - t0
print("learner: 400 iterations with collectionDuringLearn in %.1f s, last entropy %.4f, updates %d" % (dt, reps[-1]["Policy Entropy"], reps[-1]["Cumulative Model Updates"]))
PY
echo "learner rc=$?"
