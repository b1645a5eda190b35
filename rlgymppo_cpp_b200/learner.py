"""Host-side mirror of the reference's learner surface around the collection path:
``LearnerConfig`` / ``PPOLearnerConfig`` (field for field), ``ExperienceBuffer``, ``WelfordRunningStat``, ``PPOLearner``
and the ``Learner.learn()`` iteration loop.

Reference (paths under /root/reference/RLGymPPO_CPP/src/):
  public/RLGymPPO_CPP/LearnerConfig.h:14-81, PPO/PPOLearnerConfig.h:6-32, Learner.cpp:436-703,
  private/RLGymPPO_CPP/PPO/PPOLearner.cpp:67-349, PPO/ExperienceBuffer.cpp:12-121,
  public/RLGymPPO_CPP/Util/WelfordRunningStat.h:36-83.

Collection (simulation, policy/critic inference, sampling, trajectory ring, GAE, buffer rows) runs in the hand-written
CUDA engine (csrc/*.cu).  In the minibatch UPDATE the dense contractions (forward, input-gradient and weight-gradient
GEMM of every Linear layer) run on the hand-written tcgen05 TF32 kernel of csrc/gemm.cu through gemm.MLPTF32; torch
autograd strings them together and does the element-wise parts (softmax / losses / Adam / clip-by-global-norm 0.5) —
plus what the reference does not have: data-parallel replicas, one all-reduce of the flattened gradients per optimiser step (NCCL on GPUs, gloo in the
CPU tests).  PyTorch is plumbing here (memory, autograd, torch.distributed), not the product path.
"""
from __future__ import annotations

import dataclasses
import math
import time
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist

ACTION_MIN_PROB = 1e-11  # DiscretePolicy.h:19


@dataclasses.dataclass
class PPOLearnerConfig:  # PPO/PPOLearnerConfig.h:6-32
    policyLayerSizes: List[int] = dataclasses.field(default_factory=lambda: [256, 256, 256])
    criticLayerSizes: List[int] = dataclasses.field(default_factory=lambda: [256, 256, 256])
    batchSize: int = 50 * 1000
    epochs: int = 10
    policyLR: float = 3e-4
    criticLR: float = 3e-4
    entCoef: float = 0.005
    clipRange: float = 0.2
    miniBatchSize: int = 0
    autocastLearn: bool = False
    halfPrecModels: bool = False
    policyTemperature: float = 1.0
    measureGradientNoise: bool = False
    gradientNoiseUpdateInterval: int = 10
    gradientNoiseAvgDecay: float = 0.9925


@dataclasses.dataclass
class LearnerConfig:  # LearnerConfig.h:14-81 (render / metrics-sender / checkpoint fields kept for source compatibility)
    numThreads: int = 8
    numGamesPerThread: int = 16
    minInferenceSize: int = 80
    renderMode: bool = False
    renderTimeScale: float = 1.5
    renderDuringTraining: bool = False
    timestepLimit: int = 0
    expBufferSize: int = 100 * 1000
    timestepsPerIteration: int = 50 * 1000
    standardizeReturns: bool = True
    standardizeOBS: bool = False
    maxReturnsPerStatsInc: int = 150
    stepsPerObsStatsInc: int = 5
    deterministic: bool = False
    collectionDuringLearn: bool = False
    ppo: PPOLearnerConfig = dataclasses.field(default_factory=PPOLearnerConfig)
    gaeLambda: float = 0.95
    gaeGamma: float = 0.99
    rewardClipRange: float = 10.0
    checkpointLoadFolder: str = "checkpoints"
    checkpointSaveFolder: str = "checkpoints"
    saveFolderAddUnixTimestamp: bool = False
    timestepsPerSave: int = 500 * 1000
    randomSeed: int = 123
    checkpointsToKeep: int = 5
    sendMetrics: bool = True
    metricsProjectName: str = "rlgymppo-cpp"
    metricsGroupName: str = "unnamed-runs"
    metricsRunName: str = "rlgymppo-cpp-run"
    skillTrackerConfig: "object" = None  # skill_tracker.SkillTrackerConfig (LearnerConfig.h:79); None = disabled

    @property
    def num_arenas(self) -> int:
        """numThreads x numGamesPerThread Gyms (ThreadAgentManager::CreateAgents, Learner.cpp:128-133)."""
        return self.numThreads * self.numGamesPerThread


class WelfordRunningStat:
    """WelfordRunningStat.h:36-83 for shape 1 (the learner's return statistics, Learner.cpp:679-682); doubles."""

    def __init__(self):
        self.count = 0
        self.mean = 0.0
        self.var = 0.0

    def increment(self, samples, num):
        for x in np.asarray(samples[:num], dtype=np.float32).tolist():
            cur = self.count
            self.count += 1
            delta = x - self.mean
            delta_n = delta / self.count
            self.mean += delta_n
            self.var += delta * delta_n * cur

    def get_std(self) -> float:
        if self.count < 2:
            return 1.0
        v = self.var / (self.count - 1)
        if v == 0:
            v = 1.0
        return float(np.float32(math.sqrt(v)))


class ExperienceBuffer:
    """ExperienceBuffer.cpp:12-121: FIFO of maxSize rows over the tensors the PPO update reads.  Rows arrive in the
    reference's concatenation order (rlg_collector_export).  nextStates/dones/truncateds/rewards are not stored: nothing
    after GAE reads them (ExperienceBuffer.cpp:91-104 selects actions, logProbs, states, values, advantages only)."""

    KEYS = ("states", "actions", "log_probs", "values", "advantages")

    def __init__(self, max_size: int, seed: int, device):
        self.max_size = int(max_size)
        self.device = torch.device(device)
        self.cur_size = 0
        self.data: Dict[str, torch.Tensor] = {}
        self.gen = torch.Generator(device="cpu")
        self.gen.manual_seed(int(seed))

    @torch.no_grad()
    def submit(self, new: Dict[str, torch.Tensor]):
        empty = self.cur_size == 0
        first = None
        for k in self.KEYS:
            add = new[k]
            n = add.shape[0]
            first = n if first is None else first
            if n > self.max_size:
                add = add[n - self.max_size:]
                n = self.max_size
            overflow = max(self.cur_size + n - self.max_size, 0)
            start, end = self.cur_size - overflow, self.cur_size + n - overflow
            if empty:
                t = torch.empty((self.max_size,) + tuple(add.shape[1:]), dtype=add.dtype, device=self.device)
                if t.is_floating_point():
                    t.fill_(float("nan"))  # "obvious if uninitialized data is being used" (ExperienceBuffer.cpp:47-48)
                else:
                    t.zero_()
                self.data[k] = t
            elif overflow > 0:
                self.data[k][: self.cur_size - overflow] = self.data[k][overflow: self.cur_size].clone()
            self.data[k][start:end] = add
        self.cur_size = min(self.cur_size + first, self.max_size)

    def get_all_batches_shuffled(self, batch_size: int):
        """ExperienceBuffer.cpp:106-121: a fresh permutation of [0, curSize), full batches only."""
        perm = torch.randperm(self.cur_size, generator=self.gen).to(self.device)
        for start in range(0, self.cur_size - batch_size + 1, batch_size):
            idx = perm[start:start + batch_size]
            yield {k: self.data[k].index_select(0, idx) for k in self.KEYS}


def make_mlp(in_dim: int, hidden: List[int], out_dim: int) -> torch.nn.Sequential:
    """DiscretePolicy.cpp:13-27 / ValueEstimator.cpp:10-24: Linear+ReLU per hidden layer, final Linear."""
    layers, prev = [], in_dim
    for h in hidden:
        layers += [torch.nn.Linear(prev, h), torch.nn.ReLU()]
        prev = h
    layers.append(torch.nn.Linear(prev, out_dim))
    return torch.nn.Sequential(*layers)


def mlp_layers_numpy(seq: torch.nn.Sequential):
    return [(m.weight.detach().cpu().numpy(), m.bias.detach().cpu().numpy()) for m in seq if isinstance(m, torch.nn.Linear)]


class PPOLearner:
    """PPOLearner.cpp:17-349 (clipped PPO, entropy bonus, MSE value loss, clip-grad 0.5, Adam) + data-parallel replicas."""

    def __init__(self, obs_size: int, num_actions: int, cfg: PPOLearnerConfig, device, process_group=None):
        self.cfg = cfg
        self.device = torch.device(device)
        if cfg.miniBatchSize == 0:
            cfg.miniBatchSize = cfg.batchSize  # PPOLearner.cpp:19-20
        if cfg.batchSize % cfg.miniBatchSize != 0:
            raise RuntimeError("PPOLearner: batchSize must be a multiple of miniBatchSize")  # PPOLearner.cpp:22-23
        self.policy = make_mlp(obs_size, cfg.policyLayerSizes, num_actions).to(self.device)
        self.value_net = make_mlp(obs_size, cfg.criticLayerSizes, 1).to(self.device)
        self.policy_opt = torch.optim.Adam(self.policy.parameters(), lr=cfg.policyLR)
        self.value_opt = torch.optim.Adam(self.value_net.parameters(), lr=cfg.criticLR)
        # on the GPU the networks are evaluated (forward AND backward) through the hand-written tcgen05 GEMM (csrc/gemm.cu)
        # via gemm.MLPTF32, which shares these modules' parameters; the CPU path (plain torch) exists for the gloo tests of
        # the host logic only
        if self.device.type == "cuda":
            from . import gemm

            self.policy_fwd, self.value_fwd = gemm.MLPTF32(self.policy), gemm.MLPTF32(self.value_net)
        else:
            self.policy_fwd, self.value_fwd = self.policy, self.value_net
        self.pg = process_group
        self.use_cuda_graph = True
        self._graph = None
        self._graph_rows = 0
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.cumulative_model_updates = 0
        if self.world > 1:  # replicas start identical (rank 0's init)
            for p in list(self.policy.parameters()) + list(self.value_net.parameters()):
                dist.broadcast(p.data, src=0, group=self.pg)

    def update_learning_rates(self, policy_lr: float, critic_lr: float):
        """PPOLearner::UpdateLearningRates (PPOLearner.cpp:504-517).  The optimiser steps run outside the captured minibatch
        graph, so the new rates take effect at the next step without a re-capture."""
        self.cfg.policyLR, self.cfg.criticLR = float(policy_lr), float(critic_lr)
        for g in self.policy_opt.param_groups:
            g["lr"] = float(policy_lr)
        for g in self.value_opt.param_groups:
            g["lr"] = float(critic_lr)
        print(f"PPOLearner: Updated learning rate to [{policy_lr:e}, {critic_lr:e}]")

    def action_log_probs_entropy(self, obs, acts):
        """DiscretePolicy::GetBackpropData (DiscretePolicy.cpp:64-75)."""
        probs = torch.softmax(self.policy_fwd(obs) / self.cfg.policyTemperature, dim=-1).clamp(ACTION_MIN_PROB, 1)
        logp = torch.log(probs)
        return logp.gather(-1, acts.view(-1, 1).long()).view(-1), -(logp * probs).sum(-1).mean()

    def _allreduce_grads(self, module):
        if self.world == 1:
            return
        grads = [p.grad for p in module.parameters() if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg)  # ONE collective per net per optimiser step
        flat /= self.world
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def _minibatch(self, obs, acts, adv, old, tgt, acc):
        """Forward + backward of one minibatch for both networks (PPOLearner.cpp:125-271); gradients accumulate in .grad,
        diagnostics in acc (entropy, kl, ratio, value loss, clip fraction)."""
        cfg = self.cfg
        ratio_b = cfg.miniBatchSize / float(cfg.batchSize)
        vals = self.value_fwd(obs).reshape(-1)
        if cfg.policyLR != 0:
            logp, entropy = self.action_log_probs_entropy(obs, acts)
            ratio = torch.exp(logp - old)
            clipped = ratio.clamp(1 - cfg.clipRange, 1 + cfg.clipRange)
            policy_loss = -torch.min(ratio * adv, clipped * adv).mean()
            ppo_loss = (policy_loss - entropy * cfg.entCoef) * ratio_b
            with torch.no_grad():  # SB3-style diagnostics (PPOLearner.cpp:181-196)
                log_ratio = logp - old
                acc[1] += ((torch.exp(log_ratio) - 1) - log_ratio).mean()
                acc[4] += ((ratio - 1).abs() > cfg.clipRange).float().mean()
                acc[2] += ratio.mean()
                acc[0] += entropy.detach()
            ppo_loss.backward()
        if cfg.criticLR != 0:
            value_loss = torch.nn.functional.mse_loss(vals, tgt) * ratio_b
            value_loss.backward()
            acc[3] += value_loss.detach()

    def _graph_minibatch(self, mb, acc):
        """The minibatch step as ONE CUDA-graph replay (~100 small launches: the GEMMs of csrc/gemm.cu plus the element-wise
        kernels): captured once per minibatch shape, inputs and the diagnostics vector are static buffers."""
        n = mb["states"].shape[0]
        if self._graph is None or self._graph_rows != n:
            st = {k: torch.empty_like(mb[k]) for k in ExperienceBuffer.KEYS}
            st_acc = torch.zeros(5, dtype=torch.float32, device=self.device)
            for k in st:
                st[k].copy_(mb[k])
            params = list(self.policy.parameters()) + list(self.value_net.parameters())
            saved = [None if p.grad is None else p.grad.clone() for p in params]
            for p in params:  # capture with defined grads: backward then ACCUMULATES in place
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up outside the capture (lazy initialisations, cudaFuncSetAttribute)
                    self._minibatch(st["states"], st["actions"], st["advantages"], st["log_probs"], st["values"], st_acc)
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._minibatch(st["states"], st["actions"], st["advantages"], st["log_probs"], st["values"], st_acc)
            for p, g in zip(params, saved):  # the warm-up / capture must not leave a trace in the gradients
                if g is None:
                    p.grad.zero_()
                else:
                    p.grad.copy_(g)
            self._graph, self._graph_rows, self._graph_in, self._graph_acc = graph, n, st, st_acc
        for k in ExperienceBuffer.KEYS:
            self._graph_in[k].copy_(mb[k])
        self._graph_acc.zero_()
        self._graph.replay()
        acc += self._graph_acc

    def learn(self, exp: ExperienceBuffer, report: dict):
        cfg = self.cfg
        n_iter = n_mb = 0
        # diagnostics accumulate on the device and are read once at the end: no host sync inside the minibatch loop
        acc = torch.zeros(5, dtype=torch.float32, device=self.device)  # entropy, kl, ratio, value loss, clip fraction
        n_clip = 0
        before_p = torch.cat([p.detach().reshape(-1) for p in self.policy.parameters()]).clone()
        before_c = torch.cat([p.detach().reshape(-1) for p in self.value_net.parameters()]).clone()
        train_policy, train_critic = cfg.policyLR != 0, cfg.criticLR != 0
        t0 = time.perf_counter()
        for _ in range(cfg.epochs):
            for batch in exp.get_all_batches_shuffled(cfg.batchSize):
                self.policy_opt.zero_grad(set_to_none=False)
                self.value_opt.zero_grad(set_to_none=False)
                for start in range(0, cfg.batchSize, cfg.miniBatchSize):
                    stop = start + cfg.miniBatchSize
                    mb = {k: batch[k][start:stop] for k in ExperienceBuffer.KEYS}
                    if self.device.type == "cuda" and self.use_cuda_graph:
                        self._graph_minibatch(mb, acc)
                    else:
                        self._minibatch(mb["states"], mb["actions"], mb["advantages"], mb["log_probs"], mb["values"], acc)
                    n_clip += 1 if train_policy else 0
                    n_mb += 1
                if train_policy:
                    self._allreduce_grads(self.policy)
                    torch.nn.utils.clip_grad_norm_(self.policy.parameters(), 0.5)
                    self.policy_opt.step()
                if train_critic:
                    self._allreduce_grads(self.value_net)
                    torch.nn.utils.clip_grad_norm_(self.value_net.parameters(), 0.5)
                    self.value_opt.step()
                n_iter += 1
        n_iter, n_mb = max(n_iter, 1), max(n_mb, 1)
        after_p = torch.cat([p.detach().reshape(-1) for p in self.policy.parameters()])
        after_c = torch.cat([p.detach().reshape(-1) for p in self.value_net.parameters()])
        self.cumulative_model_updates += n_iter
        mean_entropy, mean_div, mean_ratio, mean_val_loss, clip_sum = (float(x) for x in acc.tolist())  # the one sync
        total = time.perf_counter() - t0
        report.update({
            "PPO Batch Consumption Time": total / n_iter, "Cumulative Model Updates": self.cumulative_model_updates,
            "Policy Entropy": mean_entropy / n_mb, "Mean KL Divergence": mean_div / n_mb, "Mean Ratio": mean_ratio / n_mb,
            "Value Function Loss": mean_val_loss / n_mb, "SB3 Clip Fraction": clip_sum / n_clip if n_clip else 0.0,
            "Policy Update Magnitude": float((before_p - after_p).norm()), "Value Function Update Magnitude": float((before_c - after_c).norm()),
            "PPO Learn Time": total,
        })


class Learner:
    """Learner.cpp:17-156 (wiring) and :436-606 (Learn loop) over the device engine.  One process per GPU; with
    torch.distributed initialised every rank owns ``cfg.num_arenas`` arenas (global ids offset by rank) and the PPO
    update is data parallel."""

    def __init__(self, engine_cfg, cfg: LearnerConfig, device_index: int = 0, iteration_callback=None, state_setter=None):
        from . import abi, collector, engine  # CUDA extension: fails loudly if missing

        self.cfg = cfg
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.device = torch.device("cuda", device_index)
        torch.manual_seed(cfg.randomSeed)  # Learner.cpp:59
        torch.backends.cuda.matmul.allow_tf32 = True
        engine_cfg.device = device_index
        engine_cfg.seed = cfg.randomSeed
        engine_cfg.arena_id_base = self.rank * engine_cfg.num_arenas
        if state_setter is not None:  # a user StateSetter on the host (collector.set_state_setter)
            engine_cfg.state_setter = abi.RLG_SETTER_HOST
        self.engine = engine.Engine(engine_cfg)
        A, P = self.engine.A, self.engine.P
        self.steps_per_iter = max(1, math.ceil(cfg.timestepsPerIteration / (self.world * A * P)))  # CollectTimesteps: >= amount rows
        self.collector = collector.Collector(self.engine, tuple(cfg.ppo.policyLayerSizes), tuple(cfg.ppo.criticLayerSizes),
                                             max_steps=self.steps_per_iter, seed=cfg.randomSeed,
                                             temperature=cfg.ppo.policyTemperature, deterministic=cfg.deterministic)
        self.ppo = PPOLearner(self.engine.obs_size, abi.RLG_NUM_ACTIONS, cfg.ppo, self.device)
        self.exp = ExperienceBuffer(max(cfg.expBufferSize // self.world, 1), cfg.randomSeed + self.rank, self.device)
        self.return_stats = WelfordRunningStat()
        self.total_timesteps = 0
        self.total_epochs = 0
        self.iteration_callback = iteration_callback
        self.skill_tracker = None
        stc = cfg.skillTrackerConfig
        if stc is not None and stc.enabled and self.rank == 0:  # eval arenas live on rank 0 only (Learner.cpp:137-144)
            from . import skill_tracker

            self.skill_tracker = skill_tracker.SkillTracker.on_engine(stc, engine_cfg, tuple(cfg.ppo.policyLayerSizes), device_index, cfg.randomSeed)
        self._push_weights()
        if state_setter is not None:
            self.collector.set_state_setter(state_setter)
            self.collector.reset_with_setter()
        else:
            self.engine.reset()

    def update_learning_rates(self, policy_lr: float, critic_lr: float):
        """Learner::UpdateLearningRates (Learner.cpp:705-707), e.g. from the iteration callback."""
        self.ppo.update_learning_rates(policy_lr, critic_lr)

    def save(self, folder=None):
        """Learner::Save (Learner.cpp:244-281), reference on-disk layout (checkpoint.py)."""
        from . import checkpoint

        return checkpoint.save_learner(self, folder)

    def load(self, folder=None):
        """Learner::Load (Learner.cpp:283-365)."""
        from . import checkpoint

        return checkpoint.load_learner(self, folder)

    def _push_weights(self):
        self.collector.set_weights(0, mlp_layers_numpy(self.ppo.policy))
        self.collector.set_weights(1, mlp_layers_numpy(self.ppo.value_net))

    def _add_new_experience(self, report: dict):
        """Learner::AddNewExperience (Learner.cpp:608-703): value preds + GAE happen on the device inside the collector."""
        cfg, col = self.cfg, self.collector
        ret_std = self.return_stats.get_std() if cfg.standardizeReturns else 1.0
        col.gae(cfg.gaeGamma, cfg.gaeLambda, ret_std, cfg.rewardClipRange)
        v = col.view()
        n = v.T * v.N
        f = lambda *shape, dt=torch.float32: torch.empty(shape, dtype=dt, device=self.device)
        new = {"states": f(n, v.obs_size), "actions": f(n, dt=torch.int64), "log_probs": f(n), "values": f(n), "advantages": f(n)}
        col.export_rows(states=new["states"].data_ptr(), actions=new["actions"].data_ptr(), log_probs=new["log_probs"].data_ptr(),
                        value_targets=new["values"].data_ptr(), advantages=new["advantages"].data_ptr())
        self.engine.sync()
        ret = col.read("ret")  # [T, N]
        returns_ref_order = np.ascontiguousarray(ret.T).reshape(-1)
        report["Avg Return"] = float(np.abs(returns_ref_order).mean()) / ret_std
        report["Avg Advantage"] = float(new["advantages"].abs().mean())
        report["Avg Val Target"] = float(new["values"].abs().mean())
        if cfg.standardizeReturns:
            stats = returns_ref_order[: cfg.maxReturnsPerStatsInc]
            if self.world > 1:  # replicas stay identical: rank 0's samples
                t = torch.from_numpy(stats.copy()).to(self.device)
                dist.broadcast(t, src=0)
                stats = t.cpu().numpy()
            self.return_stats.increment(stats, len(stats))
        self.exp.submit(new)
        return n

    def learn(self, max_iterations: Optional[int] = None):
        cfg = self.cfg
        it = 0
        reports = []
        while (cfg.timestepLimit == 0 or self.total_timesteps < cfg.timestepLimit) and (max_iterations is None or it < max_iterations):
            report = {}
            t0 = time.perf_counter()
            self.collector.collect(self.steps_per_iter)
            self.engine.sync()
            t_collect = time.perf_counter() - t0
            collected = self.steps_per_iter * self.engine.A * self.engine.P * self.world
            self.total_timesteps += collected
            if cfg.ppo.policyLR == 0 and cfg.ppo.criticLR == 0:
                it += 1
                continue
            if cfg.deterministic:
                raise RuntimeError("Learner::Learn(): Cannot run PPO learn iteration when on deterministic mode!")  # Learner.cpp:494-499
            self._add_new_experience(report)
            self.ppo.learn(self.exp, report)
            self._push_weights()
            torch.cuda.synchronize()
            if self.skill_tracker is not None:  # Learner.cpp:526-538
                self.skill_tracker.run_games(mlp_layers_numpy(self.ppo.policy), collected)
                for mode, rating in self.skill_tracker.cur_rating.items():
                    report["Skill Rating" + ("" if mode == "" else " ") + mode] = rating
            self.total_epochs += cfg.ppo.epochs
            t_total = time.perf_counter() - t0
            rew = self.collector.read("reward")
            report.update({
                "Total Iteration Time": t_total, "Collection Time": t_collect, "Consumption Time": t_total - t_collect,
                "Collected Steps/Second": int(collected / max(t_collect, 1e-9)), "Overall Steps/Second": int(collected / max(t_total, 1e-9)),
                "Timesteps Collected": collected, "Cumulative Timesteps": self.total_timesteps, "Average Step Reward": float(rew.mean()),
            })
            if self.iteration_callback:
                self.iteration_callback(self, report)
            reports.append(report)
            it += 1
        return reports
