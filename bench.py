#!/usr/bin/env python
"""bench.py — collection steps/sec of the B200 engine (and of the reference's CPU path with --impl reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 1v1 soccar, 16384 arenas per GPU, DefaultObs, examplemain rewards/terminals,
RandomState resets, tickSkip 8, placeholder mesh set v1, on-device policy (256x256x256) + critic inference.  A bench
"step" is one collect call = what one PPO iteration collects: 4 x (tcgen05 MLP inference + sampling -> fused Gym::Step ->
trajectory-ring append) over every arena of the rank + the bootstrap value pass + GAE = 4 * 16384 * 2 player-steps
(the reference counts player-timesteps: ThreadAgent.cpp:158).  Arenas shard across ranks with no collective on the data
path (weak scaling: per-GPU work fixed).

One JSON line on rank 0; see README/DESIGN.md for the meaning of `roofline`, `cpu_baseline`, `e2e`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ARENAS_PER_GPU = 16384
TEAM = 1
PADDED_OBS = False
ZERO_SUM = False
METRIC = "collection steps/sec (1v1, tickSkip=8)"
UNIT = "player-steps/s"


def workload_cfg(n_arenas, device, rank, seed=123):
    from rlgymppo_cpp_b200 import abi

    cfg = abi.default_cfg(num_arenas=n_arenas, team_size=TEAM, tick_skip=8, seed=seed)
    cfg.device = device
    cfg.arena_id_base = rank * n_arenas  # RNG streams keyed by GLOBAL arena id: results independent of the GPU count
    if PADDED_OBS:  # BASELINE configs[2]: DefaultOBSPadded(maxPlayers = 3) with slot shuffling
        cfg.obs_kind = abi.RLG_OBS_PADDED
        cfg.obs_max_players = 3
    if ZERO_SUM:    # ... + ZeroSumReward(CombinedReward(cfg-1 set), teamSpirit 0.3, opponentScale 1) (SURVEY 8d cfg 3)
        cfg.zero_sum = 1
        cfg.team_spirit = 0.3
        cfg.opponent_scale = 1.0
    return cfg


class stdout_to_stderr:
    """The reference logs through std::cout (RG_LOG); keep fd 1 clean so that rank 0 prints exactly ONE JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_baseline_sample(budget_s=15.0):
    """The reference's multithreaded collection loop (oracle/_ref) on the box's host cores, bounded sample."""
    from oracle import refsim
    from rlgymppo_cpp_b200 import abi

    if not refsim.available():
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref/librlref.so missing"}
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    T = max(1, min(cores, 256))
    G = 8
    cfg = abi.default_cfg(num_arenas=T * G, team_size=TEAM)
    with stdout_to_stderr():
        b = refsim.RefBench(cfg, T, G, 7)
        b.run(10)  # warm-up
        t = b.run(20)
        steps = int(max(20, min(12000, 20 * (budget_s * 0.6) / max(t, 1e-3))))
        t = b.run(steps)
        v = b.player_steps(steps) / t
        b.close()
    return {"value": v, "unit": UNIT, "cores": T, "kind": "reference",
            "sample": f"{T} threads x {G} gyms (1v1, same plugins), {steps} env-steps each, uniform random actions, Gym::Step + auto-reset; {t:.1f}s"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref), all host threads."""
    if rank != 0:
        return
    from oracle import refsim
    from rlgymppo_cpp_b200 import abi

    if not refsim.available() and os.path.isdir("/root/reference"):  # build the checker where the reference's sources are
        import subprocess

        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if not refsim.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/librlref.so (the reference compiled by oracle/Makefile) is not in this tree"}))
        sys.stdout.flush()
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    T = max(1, min(cores, 256))
    G = 8
    cfg = abi.default_cfg(num_arenas=T * G, team_size=TEAM)
    with stdout_to_stderr():
        b = refsim.RefBench(cfg, T, G, 7)
        b.run(160)  # one NoTouch horizon: the gyms reach their steady-state mix of resets / contacts, caches and threads are warm
        t_probe = b.run(100)
        # one bench step = one long burst (~2.5 s of every host thread; thread start-up is < 0.1 % of it), so that K steps are
        # >= 2 000 env-steps per gym for K >= 8 and the arm reads the same as one long run (cpu_baseline_sample); the whole arm is
        # bounded to ~100 s of bursts whatever K is (the default K = 500 would otherwise keep every host core busy for 20 minutes)
        burst_s = min(2.5, 100.0 / max(args.steps, 1))
        inner = int(max(50, min(20000, 100 * burst_s / max(t_probe, 1e-3))))
        for _ in range(min(args.warmup, 3)):
            b.run(inner)
        ts = [b.run(inner) for _ in range(args.steps)]
        b.close()
    total = sum(ts)
    v = T * G * inner * b.P * args.steps / total
    sample = (f"{T} threads x {G} gyms, {inner} env-steps per gym per bench step x {args.steps} steps = {inner * args.steps} env-steps per gym after a "
              f"160-step settle + {min(args.warmup, 3)} warm-up bursts (bounded sample of the 16384-arena workload); per-step rates min/median/max "
              f"{T * G * inner * b.P / max(ts) / 1e6:.3f}/{T * G * inner * b.P / float(np.median(ts)) / 1e6:.3f}/{T * G * inner * b.P / min(ts) / 1e6:.3f} M/s")
    ppo_ref = None
    if not args.no_ppo:
        try:
            ppo_ref = reference_ppo_learn_time(args)
        except Exception as ex:
            ppo_ref = {"error": f"{type(ex).__name__}: {ex}"}
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "cfg2: 1v1 soccar, 16384 arenas/GPU, DefaultObs + examplemain rewards/terminals, RandomState, tickSkip 8, "
                               "uniform random actions, placeholder mesh set v1", "reference_sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": T, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if ppo_ref is not None:
        out["ppo_iteration"] = ppo_ref
    emit(json.dumps(out))


def reference_ppo_learn_time(args, iters=4):
    """The metric's second half on the reference side: PPOLearner::Learn (PPOLearner.cpp:67-349) as the reference runs it — eager
    libtorch operators (autograd, cuBLAS GEMMs, clip_grad_norm_, Adam) — through the line-by-line torch restatement
    oracle/ppo_torch.py on THIS box's GPU (CPU when there is none), same batch shape as our arm's ppo_iteration: 131 072 rows,
    4 minibatches, 1 epoch, 256x256x256 nets.  Both cuBLAS modes are timed: fp32 (libtorch's default, what the reference gets)
    and TF32-allowed (the arithmetic our update uses)."""
    import torch

    from oracle import ppo_torch as PT
    from rlgymppo_cpp_b200 import learner as L

    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    rows = args.arenas * 2 * TEAM * args.env_steps
    if dev == "cpu":
        rows = min(rows, 8192)
    obs = 51 + 19 * 2 * TEAM
    out = {"device": dev, "rows": rows, "config": "batch = rows, 4 minibatches, 1 epoch, Adam lr 2e-4, policy/critic 256x256x256 (oracle/ppo_torch.py: eager torch)"}
    g = np.random.default_rng(0)
    data = {"states": torch.from_numpy(g.normal(size=(rows, obs)).astype(np.float32)).to(dev), "actions": torch.from_numpy(g.integers(0, 90, size=rows)).to(dev),
            "log_probs": torch.from_numpy(np.log(g.uniform(0.008, 0.014, size=rows)).astype(np.float32)).to(dev),
            "values": torch.from_numpy(g.normal(size=rows).astype(np.float32)).to(dev), "advantages": torch.from_numpy(g.normal(size=rows).astype(np.float32)).to(dev)}
    for mode, allow in (("fp32", False), ("tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = allow
        torch.backends.cudnn.allow_tf32 = allow
        torch.manual_seed(0)
        cfg = L.PPOLearnerConfig(batchSize=rows, miniBatchSize=rows // 4, epochs=1, policyLR=2e-4, criticLR=2e-4, entCoef=0.01)
        ppo = PT.TorchPPOLearner(obs, 90, cfg, dev)
        exp = PT.ExperienceBuffer(rows, 0, dev)
        exp.submit(data)
        ts = []
        for _ in range(iters + 2):
            rep = {}
            if dev != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            ppo.learn(exp, rep)
            if dev != "cpu":
                torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        out[f"learn_time_{mode}_s"] = float(np.median(ts[2:]))
    out["reference_s"] = out["learn_time_fp32_s"]
    return out


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from rlgymppo_cpp_b200 import abi, build, collector, engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    build.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    A, T = args.arenas, args.env_steps
    cfg = workload_cfg(A, local_rank, rank)
    e = engine.Engine(cfg)
    P, OBS = e.P, e.obs_size
    col = collector.Collector(e, policy_hidden=(256, 256, 256), critic_hidden=(256, 256, 256), max_steps=T, seed=123)
    col.init_default(seed=123)  # random-init weights of the examplemain architecture (256x256x256), same on every rank
    ext = torch.cuda.ExternalStream(e.stream, device=torch.device("cuda", local_rank))
    K, W = args.steps, args.warmup
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    e.reset()
    e.sync()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_iteration():
        # what one PPO iteration's collection does: T x (policy+critic inference -> fused Gym::Step -> ring append),
        # bootstrap value pass, GAE
        col.collect(T)
        col.gae(0.99, 0.95, 1.0, 10.0)

    # clocks / throttle reasons are sampled under load: from the warm-up on (nvidia-smi needs ~0.1 s to produce its first line)
    sampler = ClockSampler(local_rank)
    sampler.start()
    # settle the arenas into their steady-state mix (resets, contacts) before timing: at least 160 env-steps (one NoTouch horizon)
    W_run = max(W, -(-160 // T))
    with torch.cuda.stream(ext):
        for i in range(W_run):
            one_iteration()
    barrier()
    launches0 = e.launch_count + col.launch_count
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    step_ms = infer_ms = 0.0
    step_n = infer_n = 0
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(K):
        if not args.no_l2_flush:
            flush.fill_(i & 255)  # L2 flush between timed iterations (torch stream)
        torch.cuda.synchronize()
        with torch.cuda.stream(ext):
            starts[i].record(ext)
            one_iteration()
            ends[i].record(ext)
        torch.cuda.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = e.launch_count + col.launch_count - launches0
    # per-launch durations (roofline legs): the same steps once more with a CUDA event around every launch on the engine's stream.  Kept out
    # of the timed region above because an event between two launches serialises them, and the collector overlaps each inference with
    # the tail of the fused step before it (programmatic dependent launch + per-block ready flags, csrc/collector.cu launch_infer)
    K_t = min(K, 50)
    col.enable_timing(True)
    for i in range(K_t):
        if not args.no_l2_flush:
            flush.fill_(i & 255)
        torch.cuda.synchronize()
        with torch.cuda.stream(ext):
            one_iteration()
        sm, sn, im, inn = col.kernel_times()  # waits for this iteration's events
        step_ms += sm; step_n += sn; infer_ms += im; infer_n += inn
    col.enable_timing(False)
    barrier()
    clocks = sampler.stop()
    ms = [s.elapsed_time(t) for s, t in zip(starts, ends)]
    total_ms = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / K
    value = world * A * P * T * K / (total_ms * 1e-3)
    nb = 5 if K >= 5 else 1  # the K timed steps as 5 blocks: median block rate (this rank) next to the whole-region value
    block_rates = [world * A * P * T * len(b) / (sum(b) * 1e-3) for b in np.array_split(np.array(ms), nb) if len(b)]

    # e2e_collect: the SAME workload as `value` through the public API with HOST buffers on both sides: the policy / critic weights
    # come from page-locked host memory every step (what ThreadAgentManager::SetNewPolicy hands the agents), collect + GAE run on
    # the device, and the ExperienceBuffer rows (states, actions, log-probs, value targets, advantages) land in page-locked host memory
    rows_c = A * P * T
    st_d = torch.empty((rows_c, OBS), dtype=torch.float32, device="cuda"); ac_d = torch.empty(rows_c, dtype=torch.int64, device="cuda")
    lp_d, vt_d, ad_d = (torch.empty(rows_c, dtype=torch.float32, device="cuda") for _ in range(3))
    host_rows = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in (st_d, ac_d, lp_d, vt_d, ad_d)]
    n_ec = min(max(K, 8), 40)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(n_ec):
        col.set_weights(0, col.weights[0]); col.set_weights(1, col.weights[1])   # H2D: 1.3 MB of weights
        col.collect(T)
        col.gae(0.99, 0.95, 1.0, 10.0)
        col.export_rows(states=st_d.data_ptr(), actions=ac_d.data_ptr(), log_probs=lp_d.data_ptr(), value_targets=vt_d.data_ptr(), advantages=ad_d.data_ptr())
        e.sync()
        for h, d in zip(host_rows, (st_d, ac_d, lp_d, vt_d, ad_d)):
            h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    ec_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ec_t, op=dist.ReduceOp.MAX)
    e2e_collect = {"value": world * rows_c * n_ec / float(ec_t.item()), "unit": UNIT,
                   "h2d_bytes_per_step": int(sum(W_.nbytes + b_.nbytes for net in col.weights for W_, b_ in net)),
                   "d2h_bytes_per_step": int(sum(h.numel() * h.element_size() for h in host_rows)), "steps": n_ec,
                   "call": "Collector.set_weights (host) -> collect -> gae -> export_rows -> D2H of the ExperienceBuffer rows into page-locked memory"}

    # e2e: the reference-facing Gym::Step call with HOST buffers (H2D action indices from pinned memory, fused step,
    # D2H obs/reward/done), i.e. what a host-side policy (the reference's ThreadAgent) would drive
    n_e2e = min(max(4 * K, 16), 200)
    host_actions = np.random.default_rng(1000 + rank).integers(0, 90, size=(n_e2e + 1, A * P)).astype(np.int32)
    pinned_actions, pinned_obs, pinned_rew, pinned_done = e.host_buffers()  # the engine's page-locked host buffers
    pinned_actions[:] = host_actions[0]
    e.step_pinned()
    barrier()
    t0 = time.perf_counter()
    e2e_checksum = 0.0
    for i in range(n_e2e):
        pinned_actions[:] = host_actions[i + 1]          # this step's inputs, written by the host
        e.step_pinned()                                  # H2D actions -> fused Gym::Step -> D2H obs/reward/done, synchronous
        e2e_checksum += float(pinned_rew[::997].sum())   # the host reads the step's result
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * A * P * n_e2e / float(e2e_t.item())
    h2d = A * P * 4
    d2h = A * P * OBS * 4 + A * P * 4 + A

    ppo_iter = None
    if not args.no_ppo:
        # the metric's second half: one whole PPO iteration (collect -> GAE -> ExperienceBuffer -> minibatch update -> push
        # weights) of the examplemain-shaped learner (batch = what one iteration collects, 4 minibatches, 1 epoch)
        try:
            ppo_iter = ppo_iteration_time(args, rank, local_rank, world)
        except Exception as ex:  # reported, never required for the collection number
            ppo_iter = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        S = e.state_bytes
        alg_bytes = A * (2 * S + 4 * P + 4 * P * OBS + 4 * P + 1)
        k_step_ms = step_ms / max(step_n, 1)
        achieved = alg_bytes / (k_step_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_k_step.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        # dense part: policy fwd + critic fwd per sample, every env-step, + one critic pass for the bootstrap
        fl_pol = 2 * (OBS * 256 + 2 * 256 * 256 + 256 * 90)
        fl_cri = 2 * (OBS * 256 + 2 * 256 * 256 + 256 * 1)
        mlp_flops = A * P * (T * (fl_pol + fl_cri) + fl_cri) * K_t
        mlp_tflops = mlp_flops / (infer_ms * 1e-3) / 1e12 if infer_ms > 0 else None
        try:
            tf_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]) / 2  # TF32 = half the bf16 rate
        except Exception:
            tf_peak = 1125.0
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f32 (sim, gym layer, GAE) / tf32-in fp32-acc (MLP)", "data": "synthetic",
            "value_median_of_blocks": float(np.median(block_rates)), "block_rates": block_rates, "warmup_run": W_run,
            "config": {"workload": f"{'cfg2: ' if (TEAM == 1 and A == ARENAS_PER_GPU) else 'sweep: '}{TEAM}v{TEAM} soccar, {A} arenas/GPU, {'DefaultOBSPadded(3)' if PADDED_OBS else 'DefaultObs'} + examplemain rewards{' in ZeroSumReward(0.3)' if ZERO_SUM else ''}/terminals, RandomState, tickSkip 8, "
                                   "on-device policy (256x256x256) + critic inference, sampling, trajectory ring, GAE; placeholder mesh set v1; "
                                   f"one bench step = one collect of {T} env-steps over every arena + GAE",
                       "arenas_per_gpu": A, "players_per_arena": P, "obs_size": OBS, "tick_skip": 8, "env_steps_per_bench_step": T,
                       "player_steps_per_bench_step": world * A * P * T, "l2_flush_between_steps": not args.no_l2_flush,
                       "state_bytes_per_arena": S, "mlp_dtype": "tf32 inputs, fp32 accumulate (tcgen05)",
                       "parallelism": f"arena-sharded x{world}, no data-path collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "call": "rlg_engine_step_pinned (Gym::Step for every arena through the engine's page-locked host buffers: H2D action indices, "
                            "fused step, D2H obs/reward/done; uniform random host actions)", "reward_checksum": e2e_checksum},
            "e2e_collect": e2e_collect,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "k_roles (fused Gym::Step)", "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": k_step_ms, "launches_timed": step_n,
                         "share_of_step": (step_ms / K_t) / ms_per_step, "peak_source": peak_src,
                         "timed_in": f"a second pass of {K_t} bench steps with an event around every launch (launches serialised); in the timed region each "
                                     "inference overlaps the tail of the step before it, so the shares may add up to more than 1"},
            "roofline_mlp": {"bound": "tensor", "achieved": mlp_tflops, "peak": tf_peak, "unit": "TFLOP/s",
                             "frac": (mlp_tflops / tf_peak) if mlp_tflops else None, "kernel": "k_mlp_infer",
                             "launch_ms": infer_ms / max(infer_n, 1), "launches_timed": infer_n, "share_of_step": (infer_ms / K_t) / ms_per_step,
                             "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 rate)"},
            "wall_s_timed_region": t_wall,
        }
        if ppo_iter is not None:
            out["ppo_iteration"] = ppo_iter
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline_sample()
            except Exception as ex:  # the baseline is reported, never required for the GPU number
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {ex}"}
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_cfg4(args, rank, local_rank, world):
    """BASELINE configs[3] end to end: 3v3 soccar, a user StateSetter on the host (every reset goes through Python), ELO skill tracking
    against frozen versions on rank 0, collectionDuringLearn, the PPO update data parallel over the ranks.  One bench step = one
    Learner iteration; value = player-timesteps per second over whole iterations (collection + consumption)."""
    import torch
    import torch.distributed as dist

    from rlgymppo_cpp_b200 import build, learner, skill_tracker

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    build.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    global TEAM
    TEAM = 3
    A, T = args.arenas, args.env_steps
    rows = A * 6 * T
    cfg = learner.LearnerConfig(timestepsPerIteration=rows * world, expBufferSize=rows * world, randomSeed=123, collectionDuringLearn=True)
    cfg.sendMetrics = False
    cfg.checkpointLoadFolder = cfg.checkpointSaveFolder = ""
    cfg.ppo = learner.PPOLearnerConfig(batchSize=rows * world, miniBatchSize=rows * world // 4, epochs=1, policyLR=2e-4, criticLR=2e-4, entCoef=0.01)
    cfg.skillTrackerConfig = skill_tracker.SkillTrackerConfig(enabled=True, numEnvs=8, simTime=8 * 8 * 4 / 120, updateInterval=2, timestepsPerVersion=4 * rows * world)
    rng = np.random.default_rng(1234 + rank)

    def setter(ids, cars, balls):  # "custom StateSetter": ball dropped over midfield, cars on their own half facing it
        n, P = cars.shape
        balls["pos"][:, 0] = rng.uniform(-2000, 2000, n); balls["pos"][:, 1] = 0; balls["pos"][:, 2] = rng.uniform(300, 1200, n)
        side = np.where(cars["team"] == 0, -1.0, 1.0)
        cars["pos"][:, :, 0] = rng.uniform(-2500, 2500, (n, P)); cars["pos"][:, :, 1] = side * rng.uniform(1500, 3500, (n, P)); cars["pos"][:, :, 2] = 17
        yaw = np.where(side < 0, np.pi / 2, -np.pi / 2)
        cars["rot_forward"][:, :, 0] = np.cos(yaw); cars["rot_forward"][:, :, 1] = np.sin(yaw); cars["rot_forward"][:, :, 2] = 0
        cars["rot_right"][:, :, 0] = -np.sin(yaw); cars["rot_right"][:, :, 1] = np.cos(yaw); cars["rot_right"][:, :, 2] = 0
        cars["boost"] = 50.0

    with stdout_to_stderr():
        L = learner.Learner(workload_cfg(A, local_rank, rank), cfg, device_index=local_rank, state_setter=setter)
        K, W = args.steps, max(args.warmup, 3)
        L.learn(max_iterations=W)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        reports = L.learn(max_iterations=K)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    if rank == 0:
        med = lambda k: float(np.median([r[k] for r in reports if k in r]))
        out = {"metric": "PPO iteration, configs[3] (3v3, host StateSetter, ELO skill tracking, collection during learn)", "value": world * rows * K / dt,
               "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32 (sim) / tf32-in fp32-acc (MLP, PPO update)", "data": "synthetic",
               "config": {"workload": f"cfg4: 3v3 soccar, {A} arenas/GPU, DefaultObs, Python host StateSetter, SkillTracker (8 eval arenas on rank 0), "
                                      f"collectionDuringLearn, one bench step = one Learner iteration of {rows * world} timesteps (global batch, 4 minibatches, 1 epoch)",
                          "arenas_per_gpu": A, "players_per_arena": 6, "timesteps_per_iteration": rows * world, "parallelism": f"arena-sharded x{world}, NCCL gradient all-reduce"},
               "iteration": {"total_s": med("Total Iteration Time"), "learn_s": med("PPO Learn Time"), "entropy": med("Policy Entropy"),
                             "skill_rating": reports[-1].get("Skill Rating 3v3")},
               "gpu_launches": int(L.engine.launch_count + L.collector.launch_count + L.ppo.dev.launch_count)}
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def ppo_iteration_time(args, rank, local_rank, world, iters=6):
    """"Total Iteration Time" (Learner.cpp:558) of the device learner on the bench workload, median of the last iterations."""
    import torch

    from rlgymppo_cpp_b200 import learner

    A, T = args.arenas, args.env_steps
    rows = A * 2 * TEAM * T  # per rank
    cfg = learner.LearnerConfig(timestepsPerIteration=rows * world, expBufferSize=rows * world, randomSeed=123)
    cfg.sendMetrics = False
    cfg.checkpointLoadFolder = cfg.checkpointSaveFolder = ""
    cfg.ppo = learner.PPOLearnerConfig(batchSize=rows * world, miniBatchSize=rows * world // 4, epochs=1, policyLR=2e-4, criticLR=2e-4, entCoef=0.01)  # GLOBAL sizes
    out = {}
    for key, during in (("", False), ("collection_during_learn", True)):
        cfg.collectionDuringLearn = during
        L = learner.Learner(workload_cfg(A, local_rank, rank), cfg, device_index=local_rank)
        reports = L.learn(max_iterations=iters + (2 if during else 0))
        torch.cuda.synchronize()
        tail = reports[2:-1] if during else reports[2:]
        med = lambda k: float(np.median([r[k] for r in tail]))
        r = {"total_iteration_time_s": med("Total Iteration Time"), "collection_time_s": med("Collection Time"),
             "consumption_time_s": med("Consumption Time"), "ppo_learn_time_s": med("PPO Learn Time"), "ppo_learn_device_time_s": med("PPO Learn Device Time"),
             "timesteps_per_iteration": int(tail[-1]["Timesteps Collected"]), "overall_steps_per_s": med("Overall Steps/Second"), "iterations_timed": len(tail)}
        if key:
            out[key] = r
        else:
            out.update(r)
            out["ppo_launches_per_iteration"] = int(L.ppo.dev.launch_count // max(len(reports), 1))
        del L
    out["config"] = (f"global batch {rows * world} rows ({rows}/rank), 4 minibatches, 1 epoch, Adam lr 2e-4, policy/critic 256x256x256; the whole update on the device "
                     f"(csrc/ppo.cu: gather, tcgen05/TMA TF32 GEMMs of csrc/gemm.cu, fused loss forward+backward, clip-by-norm + Adam), {world} data-parallel rank(s), "
                     "one NCCL all-reduce of the flat gradient per optimiser step")
    return out


_REAL_STDOUT = None


def emit(line: str):
    """The ONE JSON line goes to the process's real stdout; everything else (NCCL banners, the reference's RG_LOG
    std::cout chatter) was redirected to stderr at start-up."""
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (line + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500, help="timed collects (default 500 = 2 000 env-steps per arena)")
    ap.add_argument("--warmup", type=int, default=40, help="untimed collects (floored at 160 env-steps: one NoTouch horizon)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arenas", type=int, default=ARENAS_PER_GPU, help="arenas per GPU (default: BASELINE configs[1])")
    ap.add_argument("--env-steps", type=int, default=4, help="env-steps per bench step (one collect call; cfg1's 100k timesteps/iteration ~ 4 x 32768)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ppo", action="store_true", help="skip the PPO iteration-time measurement")
    ap.add_argument("--no-l2-flush", action="store_true", help="profiling only (ncu DRAM-traffic capture of the kernel alone: the flush buffer's "
                    "dirty lines are written back during the next kernels and would be counted as theirs); a bench line needs the flush")
    ap.add_argument("--team", type=int, default=1, help="players per team (sweep only: the headline metric is quoted on 1v1)")
    ap.add_argument("--padded-obs", action="store_true", help="DefaultOBSPadded(3) (sweep only: BASELINE configs[2])")
    ap.add_argument("--zero-sum", action="store_true", help="ZeroSumReward(teamSpirit 0.3) around the reward set (sweep only)")
    ap.add_argument("--total-arenas", type=int, default=0, help="STRONG scaling: a fixed pool split evenly over the GPUs (overrides --arenas)")
    ap.add_argument("--cfg4", action="store_true", help="BASELINE configs[3]: 3v3, host StateSetter, ELO skill tracking, collection during learn, "
                    "data-parallel PPO update (NCCL all-reduce): whole Learner iterations instead of the collection loop")
    args = ap.parse_args()
    args.strong = args.total_arenas > 0
    global TEAM, METRIC, PADDED_OBS, ZERO_SUM
    PADDED_OBS, ZERO_SUM = args.padded_obs, args.zero_sum
    if args.team != 1:
        TEAM = args.team
        METRIC = f"collection steps/sec ({TEAM}v{TEAM}, tickSkip=8)"
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.total_arenas:
        if args.total_arenas % world:
            raise SystemExit("--total-arenas must be a multiple of the GPU count")
        args.arenas = args.total_arenas // world
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.cfg4:
        run_cfg4(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
