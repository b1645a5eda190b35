#!/usr/bin/env python
"""Learning-curve A/B (north_star: "learning curves statistically indistinguishable"; the reference's own correctness claim is
/root/reference/README.md:30).  TEST / EVIDENCE TOOL: the reference arm imports oracle/ (the compiled, unmodified reference).

Two arms, the SAME learner on both (rlgymppo_cpp_b200.learner: device PPO update, device GAE, same initial weights per seed, same
sampling kernel), examplemain's settings (1v1, 384 gyms, 100 k timesteps / iteration, batch 100 k, minibatch 25 k, buffer 300 k,
1 epoch, entCoef 0.01, lr 2e-4, 256x256x256):

  engine     environments = the device engine (k_roles), collection through Collector.collect
  reference  environments = oracle/_ref (the reference's Gym::Step on host threads); per env-step the observations go up, the
             tcgen05 inference kernel samples, the actions come back down; the finished trajectory is loaded into the
             collector's ring (rlg_collector_load_external) so GAE / buffer / update are the identical code

Per iteration both arms log, from the ring contents with the same host code: mean step reward, episode-end rate, goal rate
(|reward| > 25: the EventReward goal / concede term), plus the update's entropy / value loss / KL.  Output: one JSON with per-seed
curves and, per metric, mean +- SE over the seeds at matched timesteps.

    python tools/learning_curves.py --arm engine --seeds 5 --timesteps 50e6 --out gpurun_out/curves_engine.json
    python tools/learning_curves.py --arm reference --seeds 5 --timesteps 10e6 --out gpurun_out/curves_reference.json
    python tools/learning_curves.py --merge gpurun_out/curves_engine.json gpurun_out/curves_reference.json --out profiles/learning_curves_r02.json
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KEYS = ("step_reward", "episode_end_rate", "goal_rate", "entropy", "value_loss", "kl", "clip_fraction", "avg_return")


def make_cfg(seed):
    from rlgymppo_cpp_b200 import learner as L

    cfg = L.LearnerConfig(numThreads=16, numGamesPerThread=24, timestepsPerIteration=100_000, expBufferSize=300_000, randomSeed=seed)
    cfg.sendMetrics = False
    cfg.checkpointLoadFolder = cfg.checkpointSaveFolder = ""
    cfg.ppo = L.PPOLearnerConfig(batchSize=100_000, miniBatchSize=25_000, epochs=1, policyLR=2e-4, criticLR=2e-4, entCoef=0.01)
    return cfg


def ring_stats(col, report):
    rew, done = col.read("reward"), col.read("done")
    return {"step_reward": float(rew.mean()), "episode_end_rate": float(done.mean()), "goal_rate": float((np.abs(rew) > 25).mean()),
            "entropy": report["Policy Entropy"], "value_loss": report["Value Function Loss"], "kl": report["Mean KL Divergence"],
            "clip_fraction": report["SB3 Clip Fraction"], "avg_return": report["Avg Return"]}


def run_engine(seed, timesteps):
    from rlgymppo_cpp_b200 import abi, learner as L

    cfg = make_cfg(seed)
    lr = L.Learner(abi.default_cfg(num_arenas=cfg.num_arenas, team_size=1), cfg)
    curve = []
    while lr.total_timesteps < timesteps:
        rep = lr.learn(max_iterations=1)[0]
        curve.append(dict(ring_stats(lr.collector, rep), timesteps=lr.total_timesteps))
    return curve


def run_reference(seed, timesteps, threads):
    import torch

    from oracle import refsim
    from rlgymppo_cpp_b200 import abi, learner as L

    cfg = make_cfg(seed)
    ecfg = abi.default_cfg(num_arenas=cfg.num_arenas, team_size=1)
    lr = L.Learner(ecfg, cfg)  # the engine only owns the ring here: it is never stepped
    col, e = lr.collector, lr.engine
    G = cfg.num_arenas
    per = -(-G // threads)
    while G % per:
        per += 1
    pool = refsim.RefPool(abi.default_cfg(num_arenas=1, team_size=1), G // per, per, 1000 + seed)
    N, O, T = e.A * e.P, e.obs_size, lr.steps_per_iter
    assert pool.G == G and pool.obs_size == O
    obs_d = torch.empty((N, O), dtype=torch.float32, device="cuda")
    act_d = torch.empty(N, dtype=torch.int32, device="cuda")
    lp_d = torch.empty(N, dtype=torch.float32, device="cuda")
    val_d = torch.empty(N, dtype=torch.float32, device="cuda")
    obs_h = torch.empty((N, O), dtype=torch.float32).pin_memory()
    ring = dict(obs=np.zeros((T + 1, N, O), np.float32), action=np.zeros((T, N), np.int32), logprob=np.zeros((T, N), np.float32),
                reward=np.zeros((T, N), np.float32), done=np.zeros((T, e.A), np.uint8), value=np.zeros((T + 1, N), np.float32))
    obs = pool.reset()
    counter = 0
    curve = []
    while lr.total_timesteps < timesteps:
        for t in range(T + 1):
            ring["obs"][t] = obs
            obs_h.numpy()[:] = obs
            obs_d.copy_(obs_h, non_blocking=True)
            torch.cuda.synchronize()
            last = t == T
            col.infer(obs_d.data_ptr(), N, counter, 0 if last else act_d.data_ptr(), 0 if last else lp_d.data_ptr(), val_d.data_ptr())
            e.sync()
            ring["value"][t] = val_d.cpu().numpy()
            if last:
                break
            counter += 1
            ring["action"][t] = act_d.cpu().numpy()
            ring["logprob"][t] = lp_d.cpu().numpy()
            obs, rew, done = pool.step(ring["action"][t])
            ring["reward"][t] = rew
            ring["done"][t] = done
        col.load_external(ring["obs"], ring["action"], ring["logprob"], ring["reward"], ring["done"], ring["value"])
        lr.total_timesteps += T * N
        rep = {}
        lr._add_new_experience(rep)
        lr.ppo.learn(rep, e.stream)
        lr._push_weights()
        e.sync()
        curve.append(dict(ring_stats(col, rep), timesteps=lr.total_timesteps))
    pool.close()
    return curve


def summarise(curves):
    """curves: list (seeds) of lists (iterations) -> per metric mean / SE over the seeds at the common iterations."""
    n = min(len(c) for c in curves)
    out = {"timesteps": [curves[0][i]["timesteps"] for i in range(n)], "seeds": len(curves)}
    for k in KEYS:
        m = np.array([[c[i][k] for i in range(n)] for c in curves], dtype=np.float64)
        out[k] = {"mean": m.mean(0).tolist(), "se": (m.std(0, ddof=1) / np.sqrt(len(curves))).tolist() if len(curves) > 1 else [0.0] * n}
    return out


def merge(paths, out_path):
    arms = {}
    for p in paths:
        j = json.load(open(p))
        arms[j["arm"]] = j
    res = {"config": "examplemain: 1v1, 384 gyms, 100 608 timesteps / iteration, batch 100 000, minibatch 25 000, buffer 300 000, 1 epoch, "
                     "entCoef 0.01, lr 2e-4, 256x256x256; same learner (device PPO) and same initial weights per seed on both arms",
           "arms": {k: {"summary": v["summary"], "seconds": v["seconds"], "seeds": v["seeds"], "timesteps_per_seed": v["timesteps_per_seed"]} for k, v in arms.items()}}
    if "engine" in arms and "reference" in arms:
        a, b = arms["engine"]["summary"], arms["reference"]["summary"]
        n = min(len(a["timesteps"]), len(b["timesteps"]))
        cmp_ = {}
        for k in KEYS:
            am, bm = np.array(a[k]["mean"][:n]), np.array(b[k]["mean"][:n])
            se = np.hypot(np.array(a[k]["se"][:n]), np.array(b[k]["se"][:n]))
            z = (am - bm) / np.maximum(se, 1e-12)
            # windowed comparison: means over blocks of 10 iterations (the per-iteration values are noisy and auto-correlated)
            w = 10
            nb = n // w
            zb = []
            for i in range(nb):
                d = (am[i * w:(i + 1) * w] - bm[i * w:(i + 1) * w]).mean()
                s = se[i * w:(i + 1) * w].mean()
                zb.append(float(d / max(s, 1e-12)))
            cmp_[k] = {"max_abs_z_per_iteration": float(np.abs(z).max()), "frac_iterations_within_3se": float((np.abs(z) <= 3).mean()),
                       "block_z": zb, "final_engine": float(am[n - w:n].mean()), "final_reference": float(bm[n - w:n].mean())}
        res["comparison"] = {"iterations_compared": int(n), "timesteps_compared": int(a["timesteps"][n - 1]), "metrics": cmp_}
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res.get("comparison", {}), indent=1)[:3000])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["engine", "reference"])
    ap.add_argument("--seeds", type=int, default=5)
    ap.add_argument("--timesteps", type=float, default=50e6)
    ap.add_argument("--threads", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--out", required=True)
    ap.add_argument("--merge", nargs="+")
    a = ap.parse_args()
    if a.merge:
        return merge(a.merge, a.out)
    t0 = time.time()
    curves = []
    for s in range(a.seeds):
        c = run_engine(100 + s, a.timesteps) if a.arm == "engine" else run_reference(100 + s, a.timesteps, a.threads)
        curves.append(c)
        print(f"[{a.arm}] seed {s}: {len(c)} iterations, {time.time() - t0:.0f}s, last {c[-1]}", flush=True)
    with open(a.out, "w") as f:
        json.dump({"arm": a.arm, "seeds": a.seeds, "timesteps_per_seed": a.timesteps, "seconds": time.time() - t0, "summary": summarise(curves),
                   "curves": curves}, f)


if __name__ == "__main__":
    main()
