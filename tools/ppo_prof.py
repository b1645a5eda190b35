#!/usr/bin/env python
"""Where one PPO learn() goes: torch profiler over a few Learner iterations on the bench workload (tools/ppo_prof.py)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import time
import torch
from torch.profiler import profile, ProfilerActivity
from rlgymppo_cpp_b200 import abi, learner

A, T = 16384, 4
rows = A * 2 * T
cfg = learner.LearnerConfig(timestepsPerIteration=rows, expBufferSize=rows, randomSeed=123)
cfg.ppo = learner.PPOLearnerConfig(batchSize=rows, miniBatchSize=rows // 4, epochs=1, policyLR=2e-4, criticLR=2e-4, entCoef=0.01)
L = learner.Learner(abi.default_cfg(num_arenas=A, team_size=1), cfg)
L.learn(max_iterations=3)
torch.cuda.synchronize()
for name in ("graph", "nograph"):
    L.ppo.use_cuda_graph = name == "graph"
    rep = {}
    t0 = time.perf_counter(); L.ppo.learn(L.exp, rep); torch.cuda.synchronize(); t1 = time.perf_counter()
    t0 = time.perf_counter(); L.ppo.learn(L.exp, rep); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(name, "learn() wall %.2f ms" % ((t1 - t0) * 1e3))
L.ppo.use_cuda_graph = False
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    L.ppo.learn(L.exp, {})
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
