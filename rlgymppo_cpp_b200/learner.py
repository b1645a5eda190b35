"""Host-side mirror of the reference's learner surface around the collection path:
``LearnerConfig`` / ``PPOLearnerConfig`` (field for field), ``WelfordRunningStat``, ``PPOLearner`` (a thin host object over the
device learner of csrc/ppo.cu) and the ``Learner.learn()`` iteration loop.

Reference (paths under /root/reference/RLGymPPO_CPP/src/):
  public/RLGymPPO_CPP/LearnerConfig.h:14-81, PPO/PPOLearnerConfig.h:6-32, Learner.cpp:436-703,
  private/RLGymPPO_CPP/PPO/PPOLearner.cpp:67-349, PPO/ExperienceBuffer.cpp:12-121,
  public/RLGymPPO_CPP/Util/WelfordRunningStat.h:36-83.

Collection (simulation, policy/critic inference, sampling, trajectory ring, GAE, buffer rows) AND the PPO update
(experience FIFO, shuffled batches, every Linear layer's forward / input-gradient / weight-gradient GEMM on the tcgen05 TF32
kernel, the fused loss forward+backward, clip-by-global-norm + Adam) run in the hand-written CUDA library (csrc/*.cu); this
file is host logic only.  What the reference does not have: data-parallel replicas — ONE all-reduce of the flat gradient
vector of both networks per optimiser step (NCCL through torch.distributed on the device learner's stream).  PyTorch is
plumbing here (torch.distributed, checkpoint files), not the product path; there is no CPU path (the torch restatement of
the update lives in oracle/ppo_torch.py, test infrastructure).
"""
from __future__ import annotations

import dataclasses
import math
import os
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


def make_mlp(in_dim: int, hidden: List[int], out_dim: int) -> torch.nn.Sequential:
    """DiscretePolicy.cpp:13-27 / ValueEstimator.cpp:10-24: Linear+ReLU per hidden layer, final Linear (host mirror: initialisation
    stream and checkpoint files)."""
    layers, prev = [], in_dim
    for h in hidden:
        layers += [torch.nn.Linear(prev, h), torch.nn.ReLU()]
        prev = h
    layers.append(torch.nn.Linear(prev, out_dim))
    return torch.nn.Sequential(*layers)


def mlp_layers_numpy(seq: torch.nn.Sequential):
    return [(m.weight.detach().cpu().numpy(), m.bias.detach().cpu().numpy()) for m in seq if isinstance(m, torch.nn.Linear)]


def shard_sizes(cfg: LearnerConfig, world: int) -> Dict[str, int]:
    """Data-parallel replicas split the reference's GLOBAL sizes: expBufferSize, timestepsPerIteration, ppo.batchSize and
    ppo.miniBatchSize are all divided by the replica count (a replica that held fewer rows than its batch would never step)."""
    mbs = cfg.ppo.miniBatchSize or cfg.ppo.batchSize  # PPOLearner.cpp:19-20
    for name, v in (("ppo.batchSize", cfg.ppo.batchSize), ("ppo.miniBatchSize", mbs), ("expBufferSize", cfg.expBufferSize)):
        if v % world != 0:
            raise RuntimeError(f"Learner: {name} = {v} must be a multiple of the {world} data-parallel replicas")
    out = {"batchSize": cfg.ppo.batchSize // world, "miniBatchSize": mbs // world, "expBufferSize": cfg.expBufferSize // world}
    if out["batchSize"] % out["miniBatchSize"] != 0:
        raise RuntimeError("PPOLearner: batchSize must be a multiple of miniBatchSize")  # PPOLearner.cpp:22-23
    if out["expBufferSize"] < out["batchSize"]:
        raise RuntimeError(f"Learner: expBufferSize ({cfg.expBufferSize}) is smaller than ppo.batchSize ({cfg.ppo.batchSize}): no batch would ever be formed")
    return out


class _AdamStateAdapter:
    """state_dict / load_state_dict in torch.optim.Adam's format over the device learner's flat Adam moments (checkpoint files)."""

    def __init__(self, owner: "PPOLearner", net: int):
        self.owner, self.net = owner, net

    def state_dict(self):
        from . import ppo as P

        dev = self.owner.dev
        m, v = dev.get_layers(self.net, P.EXP_AVG), dev.get_layers(self.net, P.EXP_AVG_SQ)
        step = dev.adam_steps()[self.net]
        state, i = {}, 0
        for (mw, mb), (vw, vb) in zip(m, v):
            for a, b in ((mw, vw), (mb, vb)):
                state[i] = {"step": torch.tensor(float(step)), "exp_avg": torch.from_numpy(a.copy()), "exp_avg_sq": torch.from_numpy(b.copy())}
                i += 1
        lr = self.owner.cfg.policyLR if self.net == 0 else self.owner.cfg.criticLR
        group = {"lr": lr, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0, "amsgrad": False, "maximize": False, "foreach": None,
                 "capturable": False, "differentiable": False, "fused": None, "params": list(range(i))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        from . import ppo as P

        dev = self.owner.dev
        st = sd.get("state", {})
        if not st:
            return
        keys = sorted(st.keys())
        dims = dev.dims[self.net]
        if len(keys) != 2 * len(dims):
            raise RuntimeError("optimizer state does not match the network")
        m, v = [], []
        for l in range(len(dims)):
            w, b = st[keys[2 * l]], st[keys[2 * l + 1]]
            m.append((w["exp_avg"].cpu().numpy(), b["exp_avg"].cpu().numpy()))
            v.append((w["exp_avg_sq"].cpu().numpy(), b["exp_avg_sq"].cpu().numpy()))
        dev.set_layers(self.net, m, P.EXP_AVG)
        dev.set_layers(self.net, v, P.EXP_AVG_SQ)
        steps = list(dev.adam_steps())
        steps[self.net] = int(float(st[keys[0]]["step"]))
        dev.set_adam_steps(*steps)


class PPOLearner:
    """PPOLearner.cpp:17-349 on the device (csrc/ppo.cu through ppo.DevicePPO) + data-parallel replicas.  ``cfg`` carries this
    replica's batch sizes.  ``policy`` / ``value_net`` are host torch mirrors of the device parameters (initialisation from
    torch's seeded stream like the reference's libtorch modules, checkpoint files, skill-tracker snapshots): reading them
    downloads the current device weights, ``upload_modules()`` sends edited mirrors back."""

    def __init__(self, obs_size: int, num_actions: int, cfg: PPOLearnerConfig, device, exp_buffer_size: int = 100_000, seed: int = 0,
                 process_group=None):
        from . import ppo as P  # CUDA extension: fails loudly if missing

        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PPOLearner runs on a CUDA device only (csrc/ppo.cu); there is no CPU path")
        if cfg.miniBatchSize == 0:
            cfg.miniBatchSize = cfg.batchSize  # PPOLearner.cpp:19-20
        if cfg.batchSize % cfg.miniBatchSize != 0:
            raise RuntimeError("PPOLearner: batchSize must be a multiple of miniBatchSize")  # PPOLearner.cpp:22-23
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        index = self.device.index if self.device.index is not None else 0
        self.dev = P.DevicePPO(obs_size, num_actions, cfg.policyLayerSizes, cfg.criticLayerSizes, cfg.batchSize, cfg.miniBatchSize, cfg.epochs,
                               cfg.policyLR, cfg.criticLR, cfg.entCoef, cfg.clipRange, cfg.policyTemperature, exp_buffer_size, seed, index, self.world)
        self._policy = make_mlp(obs_size, cfg.policyLayerSizes, num_actions)
        self._value_net = make_mlp(obs_size, cfg.criticLayerSizes, 1)
        self._mirror_stale = False
        self.policy_opt, self.value_opt = _AdamStateAdapter(self, 0), _AdamStateAdapter(self, 1)
        self.cumulative_model_updates = 0
        self.upload_modules()
        if self.world > 1:
            ptr, count, _ = self.dev.flat(P.PARAMS)
            params = torch.as_tensor(P._DevArray(ptr, count), device=self.device)
            torch.cuda.synchronize(self.device)
            dist.broadcast(params, src=0, group=self.pg)  # replicas start identical (rank 0's init)
            torch.cuda.synchronize(self.device)
            self._refresh_after_raw_param_write()
            gptr, gcount, _ = self.dev.flat(P.GRADS)
            self._grads = torch.as_tensor(P._DevArray(gptr, gcount), device=self.device)

            def allreduce(_ptr, _count, stream):  # ONE collective for both networks per optimiser step, on the learner's stream
                with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=self.device)):
                    dist.all_reduce(self._grads, op=dist.ReduceOp.SUM, group=self.pg)

            self.dev.set_allreduce(allreduce, self.world)

    def _refresh_after_raw_param_write(self):
        self.dev.set_layers(0, self.dev.get_layers(0))  # re-derives the transposed weight copies
        self.dev.set_layers(1, self.dev.get_layers(1))
        self._mirror_stale = True

    def _download(self):
        if self._mirror_stale:
            with torch.no_grad():
                for seq, net in ((self._policy, 0), (self._value_net, 1)):
                    lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
                    for m, (W, b) in zip(lin, self.dev.get_layers(net)):
                        m.weight.copy_(torch.from_numpy(W))
                        m.bias.copy_(torch.from_numpy(b))
            self._mirror_stale = False

    @property
    def policy(self) -> torch.nn.Sequential:
        self._download()
        return self._policy

    @property
    def value_net(self) -> torch.nn.Sequential:
        self._download()
        return self._value_net

    def upload_modules(self):
        """Host mirrors -> device parameters (after initialisation or a checkpoint load)."""
        self.dev.set_layers(0, mlp_layers_numpy(self._policy))
        self.dev.set_layers(1, mlp_layers_numpy(self._value_net))
        self._mirror_stale = False

    def update_learning_rates(self, policy_lr: float, critic_lr: float):
        """PPOLearner::UpdateLearningRates (PPOLearner.cpp:504-517); a rate of 0 freezes that network from the next Learn on."""
        self.cfg.policyLR, self.cfg.criticLR = float(policy_lr), float(critic_lr)
        self.dev.set_lr(policy_lr, critic_lr)
        print(f"PPOLearner: Updated learning rate to [{policy_lr:e}, {critic_lr:e}]")

    def learn(self, report: dict, stream: int = 0):
        t0 = time.perf_counter()
        rep = self.dev.learn(stream)
        total = time.perf_counter() - t0
        self._mirror_stale = True
        if rep.batches == 0:
            print(f"PPOLearner::Learn(): WARNING: the experience buffer holds {self.dev.buffer_size} rows, fewer than batchSize "
                  f"{self.cfg.batchSize}: no optimiser step was taken", flush=True)
        self.cumulative_model_updates += int(rep.batches)
        n_iter = max(int(rep.batches), 1)
        report.update({
            "PPO Batch Consumption Time": total / n_iter, "Cumulative Model Updates": self.cumulative_model_updates,
            "Policy Entropy": rep.entropy, "Mean KL Divergence": rep.kl, "Mean Ratio": rep.ratio, "Value Function Loss": rep.value_loss,
            "SB3 Clip Fraction": rep.clip_fraction, "Policy Update Magnitude": rep.policy_update_magnitude,
            "Value Function Update Magnitude": rep.critic_update_magnitude, "PPO Learn Time": total, "PPO Learn Device Time": rep.device_ms * 1e-3,
        })
        return rep


class Learner:
    """Learner.cpp:17-156 (wiring) and :436-606 (Learn loop) over the device engine.  One process per GPU; with
    torch.distributed initialised every rank owns ``engine_cfg.num_arenas`` arenas (global ids offset by rank) and the PPO
    update is data parallel: ``expBufferSize``, ``timestepsPerIteration``, ``ppo.batchSize`` and ``ppo.miniBatchSize`` are
    GLOBAL sizes split evenly over the ranks (shard_sizes)."""

    def __init__(self, engine_cfg, cfg: LearnerConfig, device_index: int = 0, iteration_callback=None, state_setter=None, mesh_blobs=None,
                 collision_meshes_folder: Optional[str] = None, step_callback=None, action_table=None):
        from . import abi, collector, engine  # CUDA extension: fails loudly if missing

        self.cfg = cfg
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.device = torch.device("cuda", device_index)
        torch.manual_seed(cfg.randomSeed)  # Learner.cpp:59
        if engine_cfg.num_arenas != cfg.num_arenas:
            # the reference creates numThreads * numGamesPerThread Gyms (Learner.cpp:128-133); here the pool size is the engine's
            print(f"Learner: engine pool of {engine_cfg.num_arenas} arenas per GPU overrides numThreads * numGamesPerThread = {cfg.num_arenas}")
        engine_cfg.device = device_index
        engine_cfg.seed = cfg.randomSeed
        engine_cfg.arena_id_base = self.rank * engine_cfg.num_arenas
        if state_setter is not None:  # a user StateSetter on the host (collector.set_state_setter)
            engine_cfg.state_setter = abi.RLG_SETTER_HOST
        if collision_meshes_folder is not None:  # RocketSim::Init(folder): the real arena meshes (R/RocketSim.cpp:70-212)
            from . import meshes

            mesh_blobs = meshes.read_cmf_folder(collision_meshes_folder)
        self.mesh_blobs = mesh_blobs
        self.engine = engine.Engine(engine_cfg, mesh_blobs=mesh_blobs)
        if action_table is not None:  # a user ActionParser as its [n_actions, 8] table (ActionParser.h:11-14); the policy head follows it
            self.engine.set_action_table(action_table)
        A, P = self.engine.A, self.engine.P
        self.steps_per_iter = max(1, math.ceil(cfg.timestepsPerIteration / (self.world * A * P)))  # CollectTimesteps: >= amount rows
        self.collector = collector.Collector(self.engine, tuple(cfg.ppo.policyLayerSizes), tuple(cfg.ppo.criticLayerSizes),
                                             max_steps=self.steps_per_iter, seed=cfg.randomSeed,
                                             temperature=cfg.ppo.policyTemperature, deterministic=cfg.deterministic)
        sizes = shard_sizes(cfg, self.world)
        self.rank_ppo_cfg = dataclasses.replace(cfg.ppo, batchSize=sizes["batchSize"], miniBatchSize=sizes["miniBatchSize"])
        self.ppo = PPOLearner(self.engine.obs_size, self.engine.num_actions, self.rank_ppo_cfg, self.device, exp_buffer_size=sizes["expBufferSize"],
                              seed=cfg.randomSeed + self.rank)
        self.return_stats = WelfordRunningStat()
        self.total_timesteps = 0
        self.total_epochs = 0
        self.iteration_callback = iteration_callback
        self.step_callback = step_callback
        self.metric_sender = None
        self.skill_tracker = None
        stc = cfg.skillTrackerConfig
        if stc is not None and stc.enabled and self.rank == 0:  # eval arenas live on rank 0 only (Learner.cpp:137-144)
            from . import skill_tracker

            self.skill_tracker = skill_tracker.SkillTracker.on_engine(stc, engine_cfg, tuple(cfg.ppo.policyLayerSizes), device_index, cfg.randomSeed,
                                                                      mesh_blobs=mesh_blobs, action_table=action_table)
        self.save_folder = cfg.checkpointSaveFolder
        if self.save_folder and cfg.saveFolderAddUnixTimestamp:  # Learner.cpp:30-31
            self.save_folder = os.path.join(self.save_folder, str(int(time.time())))
        self.run_id = None
        if cfg.checkpointLoadFolder and os.path.isdir(cfg.checkpointLoadFolder):  # Learner.cpp:130-131 (Load() in the constructor)
            self.load()
        if cfg.sendMetrics and self.rank == 0:  # Learner.cpp:146-151
            from . import sinks

            try:
                self.metric_sender = sinks.MetricSender(cfg.metricsProjectName, cfg.metricsGroupName, cfg.metricsRunName, self.run_id)
                self.run_id = getattr(self.metric_sender, "run_id", self.run_id)
            except Exception as ex:  # the reference aborts when its python bridge is missing; here metrics are optional
                print(f"Learner: metrics disabled ({ex})")
        self._push_weights()
        if state_setter is not None:
            self.collector.set_state_setter(state_setter)
            self.collector.reset_with_setter()
        else:
            self.engine.reset()
        self._ahead = False  # collectionDuringLearn: the next iteration's collect is already queued

    def update_learning_rates(self, policy_lr: float, critic_lr: float):
        """Learner::UpdateLearningRates (Learner.cpp:705-707), e.g. from the iteration callback."""
        self.cfg.ppo.policyLR, self.cfg.ppo.criticLR = float(policy_lr), float(critic_lr)
        self.ppo.update_learning_rates(policy_lr, critic_lr)

    def save(self, folder=None):
        """Learner::Save (Learner.cpp:244-281), reference on-disk layout (checkpoint.py)."""
        from . import checkpoint

        return checkpoint.save_learner(self, folder if folder is not None else self.save_folder)

    def load(self, folder=None):
        """Learner::Load (Learner.cpp:283-365)."""
        from . import checkpoint

        return checkpoint.load_learner(self, folder)

    def _push_weights(self):
        """New policy / critic to the agents, device to device (ThreadAgentManager::SetNewPolicy)."""
        self.ppo.dev.push_weights(self.collector, self.engine.stream)

    def _add_new_experience(self, report: dict):
        """Learner::AddNewExperience (Learner.cpp:608-703): value preds + GAE happen on the device inside the collector, the
        report's means and the Welford samples are reduced there too; the rows go straight into the device FIFO."""
        cfg, col = self.cfg, self.collector
        ret_std = self.return_stats.get_std() if cfg.standardizeReturns else 1.0
        self.last_return_std = ret_std
        col.gae(cfg.gaeGamma, cfg.gaeLambda, ret_std, cfg.rewardClipRange)
        v = col.view()
        n = v.T * v.N
        means, first = col.return_stats(cfg.maxReturnsPerStatsInc if cfg.standardizeReturns else 0)
        report["Avg Return"] = means[0] / ret_std
        report["Avg Advantage"] = means[1]
        report["Avg Val Target"] = means[2]
        if cfg.standardizeReturns:
            stats = first
            if self.world > 1:  # replicas stay identical: rank 0's samples
                t = torch.from_numpy(stats.copy()).to(self.device)
                dist.broadcast(t, src=0)
                stats = t.cpu().numpy()
            self.return_stats.increment(stats, len(stats))
        self.ppo.dev.submit_collector(col, self.engine.stream)
        return n

    def learn(self, max_iterations: Optional[int] = None):
        cfg = self.cfg
        it = 0
        reports = []
        ts_since_save = 0
        es = self.engine.stream
        while (cfg.timestepLimit == 0 or self.total_timesteps < cfg.timestepLimit) and (max_iterations is None or it < max_iterations):
            report = {}
            t0 = time.perf_counter()
            if not self._ahead:
                self.collector.collect(self.steps_per_iter)
            self._ahead = False
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
            metrics = self.engine.metrics()  # ThreadAgentManager::GetMetrics of THIS iteration's steps
            self.engine.reset_metrics()
            overlap = cfg.collectionDuringLearn and (max_iterations is None or it + 1 < max_iterations)
            if overlap:
                # Learner.cpp:473-474: the agents keep collecting (with the policy they have) while the update runs.  Here both are
                # GPU work: the next iteration's collect is queued on the engine's stream, the update on the learner's own
                # stream behind the export, and the two share the SMs.
                self.engine.sync()
                self.collector.collect(self.steps_per_iter)
                self.ppo.learn(report, self.ppo.dev.stream)
                self._ahead = True
            else:
                self.ppo.learn(report, es)
            self._push_weights()  # ordered after the queued collect on the engine's stream
            if not overlap:
                self.engine.sync()
            if self.skill_tracker is not None:  # Learner.cpp:526-538
                self.skill_tracker.run_games(mlp_layers_numpy(self.ppo.policy), collected)
                for mode, rating in self.skill_tracker.cur_rating.items():
                    report["Skill Rating" + ("" if mode == "" else " ") + mode] = rating
            self.total_epochs += cfg.ppo.epochs
            t_total = time.perf_counter() - t0
            report.update({
                "Total Iteration Time": t_total, "Collection Time": t_collect, "Consumption Time": t_total - t_collect,
                "Collected Steps/Second": int(collected / max(t_collect, 1e-9)), "Overall Steps/Second": int(collected / max(t_total, 1e-9)),
                "Timesteps Collected": collected, "Cumulative Timesteps": self.total_timesteps,
                "Average Step Reward": metrics["avg_step_reward"], "Average Episode Reward": metrics["avg_episode_reward"],
            })
            if self.iteration_callback:
                self.iteration_callback(self, report)
            if self.metric_sender is not None:
                self.metric_sender.send(report)
            ts_since_save += collected
            if ts_since_save > cfg.timestepsPerSave and self.save_folder and self.rank == 0:  # Learner.cpp:585-589
                self.save()
                ts_since_save = 0
            reports.append(report)
            it += 1
        if self._ahead:
            self.engine.sync()
        return reports
