"""TEST INFRASTRUCTURE — a plain-torch fp32 restatement of the reference's PPO update, used only as the checker for the device
learner (rlgymppo_cpp_b200/ppo.py -> csrc/ppo.cu) in tests/ and as the `--impl reference` PPO leg of bench.py.  Nothing under
rlgymppo_cpp_b200/ imports this module.

Follows /root/reference/RLGymPPO_CPP/src/private/RLGymPPO_CPP/PPO/PPOLearner.cpp:67-349 (Learn), PPO/DiscretePolicy.cpp:64-75
(GetBackpropData), PPO/ExperienceBuffer.cpp:12-121 line by line with torch autograd, torch.optim.Adam and
torch.nn.utils.clip_grad_norm_ — the same libtorch operators the reference calls.  Parity pinning: the forward / log-prob /
entropy maths is pinned against the reference's compiled binaries through oracle/ppo_oracle.py (tests/test_ppo_oracle.py);
Adam and clip_grad_norm_ are torch's own on both sides (unpinned by construction).
"""
from __future__ import annotations

import time
from typing import Dict, List

import torch
import torch.distributed as dist

ACTION_MIN_PROB = 1e-11  # DiscretePolicy.h:19


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
        forced = getattr(self, "forced_perms", None)  # tests replay the device learner's permutations
        perm = (torch.as_tensor(forced.pop(0), dtype=torch.int64) if forced else torch.randperm(self.cur_size, generator=self.gen)).to(self.device)
        for start in range(0, self.cur_size - batch_size + 1, batch_size):
            idx = perm[start:start + batch_size]
            yield {k: self.data[k].index_select(0, idx) for k in self.KEYS}


def tf32_trunc(t: torch.Tensor) -> torch.Tensor:
    """fp32 -> the TF32 value a tcgen05 kind::tf32 MMA sees when it is fed raw fp32 bits: the low 13 mantissa bits are ignored."""
    return (t.contiguous().view(torch.int32) & -8192).view(torch.float32)


class _TF32LinearFn(torch.autograd.Function):
    """y = x W^T + b with both operands of EVERY contraction (forward, input gradient, weight gradient) truncated to TF32 and the
    products summed in fp32 — the arithmetic of csrc/gemm.cu's TMA path up to summation order.  With it the restatement's ReLU masks
    and clip decisions coincide with the device's, so gradients can be compared at 1e-4 instead of the ~2 % element noise that a
    plain fp32 reference shows (a hidden unit whose pre-activation is within TF32 rounding of 0 switches its whole row term)."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return tf32_trunc(x) @ tf32_trunc(w).t() + b

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dyt = tf32_trunc(dy)
        return dyt @ tf32_trunc(w), dyt.t() @ tf32_trunc(x), dy.sum(0)


class TF32Linear(torch.nn.Linear):
    def forward(self, x):
        return _TF32LinearFn.apply(x, self.weight, self.bias)


def make_mlp(in_dim: int, hidden: List[int], out_dim: int, emulate_tf32: bool = False) -> torch.nn.Sequential:
    """DiscretePolicy.cpp:13-27 / ValueEstimator.cpp:10-24: Linear+ReLU per hidden layer, final Linear."""
    lin = TF32Linear if emulate_tf32 else torch.nn.Linear
    layers, prev = [], in_dim
    for h in hidden:
        layers += [lin(prev, h), torch.nn.ReLU()]
        prev = h
    layers.append(lin(prev, out_dim))
    return torch.nn.Sequential(*layers)


def mlp_layers_numpy(seq: torch.nn.Sequential):
    return [(m.weight.detach().cpu().numpy(), m.bias.detach().cpu().numpy()) for m in seq if isinstance(m, torch.nn.Linear)]


class TorchPPOLearner:
    """PPOLearner.cpp:17-349 (clipped PPO, entropy bonus, MSE value loss, clip-grad 0.5, Adam) + data-parallel replicas."""

    def __init__(self, obs_size: int, num_actions: int, cfg: PPOLearnerConfig, device, process_group=None, emulate_tf32: bool = False):
        self.cfg = cfg
        self.device = torch.device(device)
        if cfg.miniBatchSize == 0:
            cfg.miniBatchSize = cfg.batchSize  # PPOLearner.cpp:19-20
        if cfg.batchSize % cfg.miniBatchSize != 0:
            raise RuntimeError("PPOLearner: batchSize must be a multiple of miniBatchSize")  # PPOLearner.cpp:22-23
        self.policy = make_mlp(obs_size, cfg.policyLayerSizes, num_actions, emulate_tf32).to(self.device)
        self.value_net = make_mlp(obs_size, cfg.criticLayerSizes, 1, emulate_tf32).to(self.device)
        self.policy_opt = torch.optim.Adam(self.policy.parameters(), lr=cfg.policyLR)
        self.value_opt = torch.optim.Adam(self.value_net.parameters(), lr=cfg.criticLR)
        self.policy_fwd, self.value_fwd = self.policy, self.value_net
        self.pg = process_group
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


