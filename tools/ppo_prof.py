#!/usr/bin/env python
"""One PPOLearner::Learn of the device learner (csrc/ppo.cu) on the bench shape, standalone: 131 072 synthetic rows, 4 minibatches,
256x256x256 nets.  Prints the CUDA-event time of learn(); run it under
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ppo.csv python tools/ppo_prof.py --iters 2
for the per-kernel launch list (tools/launch_shares.py summarises it)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from rlgymppo_cpp_b200 import learner as L

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=131072)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
rows, obs = a.rows, 89
torch.manual_seed(0)
cfg = L.PPOLearnerConfig(batchSize=rows, miniBatchSize=rows // 4, epochs=1, policyLR=2e-4, criticLR=2e-4, entCoef=0.01)
ppo = L.PPOLearner(obs, 90, cfg, "cuda:0", exp_buffer_size=rows, seed=1)
g = np.random.default_rng(0)
t = {"states": torch.from_numpy(g.normal(size=(rows, obs)).astype(np.float32)).cuda(), "actions": torch.from_numpy(g.integers(0, 90, size=rows)).cuda(),
     "log_probs": torch.from_numpy(np.log(g.uniform(0.008, 0.014, size=rows)).astype(np.float32)).cuda(),
     "values": torch.from_numpy(g.normal(size=rows).astype(np.float32)).cuda(), "advantages": torch.from_numpy(g.normal(size=rows).astype(np.float32)).cuda()}
torch.cuda.synchronize()
ppo.dev.submit(t["states"].data_ptr(), t["actions"].data_ptr(), t["log_probs"].data_ptr(), t["values"].data_ptr(), t["advantages"].data_ptr(), rows)
torch.cuda.synchronize()
ms = []
for i in range(a.iters):
    rep = {}
    r = ppo.learn(rep)
    ms.append(r.device_ms)
print("learn device ms:", [round(x, 3) for x in ms], "median", round(float(np.median(ms[1:] or ms)), 3), "launches/learn", ppo.dev.launch_count // a.iters)
