"""CPU suite: host-side learner logic (no engine, no device): the torch restatement of the PPO update (oracle/ppo_torch.py, the
checker the GPU tests compare csrc/ppo.cu with) — its ExperienceBuffer FIFO vs the numpy oracle, its loss/step vs an independent
second restatement of PPOLearner.cpp:125-290, the data-parallel gradient averaging with world_size 2 over gloo — plus the
product's host logic: Welford return statistics, the data-parallel size split, and that the device learner refuses a CPU."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ppo_oracle as po
from oracle import ppo_torch as PT
from rlgymppo_cpp_b200 import learner as L


def _fake_rows(n, obs, seed):
    g = np.random.default_rng(seed)
    return {"states": torch.from_numpy(g.normal(size=(n, obs)).astype(np.float32)), "actions": torch.from_numpy(g.integers(0, 90, size=n)),
            "log_probs": torch.from_numpy(np.log(g.uniform(0.005, 0.05, size=n)).astype(np.float32)),
            "values": torch.from_numpy(g.normal(size=n).astype(np.float32)), "advantages": torch.from_numpy(g.normal(size=n).astype(np.float32))}


def test_experience_buffer_matches_reference_fifo():
    b = PT.ExperienceBuffer(10, 0, "cpu")
    o = po.ExperienceBufferOracle(10)
    for i, n in enumerate([4, 4, 5, 13, 2]):
        rows = _fake_rows(n, 3, i)
        b.submit(rows)
        o.submit({k: v.numpy() for k, v in rows.items()})
        assert b.cur_size == o.cur
        for k in PT.ExperienceBuffer.KEYS:
            assert np.array_equal(b.data[k][: b.cur_size].numpy(), o.data[k][: o.cur]), (i, k)
    batches = list(b.get_all_batches_shuffled(4))
    assert len(batches) == 2  # full batches only (ExperienceBuffer.cpp:114)
    seen = torch.cat([x["values"] for x in batches])
    assert len(set(seen.tolist())) == 8 and set(seen.tolist()) <= set(b.data["values"].tolist())


def test_welford_matches_numpy():
    w = L.WelfordRunningStat()
    assert w.get_std() == 1.0
    x = np.random.default_rng(0).normal(3, 2, size=500).astype(np.float32)
    w.increment(x, 150)
    assert abs(w.get_std() - np.std(x[:150].astype(np.float64), ddof=1)) < 1e-5


def _restated_losses(policy, value_net, batch, cfg, ratio_b):
    """PPOLearner.cpp:139-178 written out independently (log-softmax form is NOT used on purpose: the reference
    clamps probabilities)."""
    logits = policy(batch["states"]) / cfg.policyTemperature
    p = torch.exp(logits - torch.logsumexp(logits, dim=-1, keepdim=True)).clamp(1e-11, 1)
    lp_all = torch.log(p)
    lp = lp_all[torch.arange(len(p)), batch["actions"]]
    ent = -(lp_all * p).sum(-1).mean()
    r = torch.exp(lp - batch["log_probs"])
    s1, s2 = r * batch["advantages"], r.clamp(1 - cfg.clipRange, 1 + cfg.clipRange) * batch["advantages"]
    ppo = (-(torch.minimum(s1, s2)).mean() - ent * cfg.entCoef) * ratio_b
    v = ((value_net(batch["states"]).view(-1) - batch["values"]) ** 2).mean() * ratio_b
    return ppo, v


def test_ppo_learn_step_matches_restatement():
    torch.manual_seed(0)
    cfg = L.PPOLearnerConfig(policyLayerSizes=[32, 32], criticLayerSizes=[32, 32], batchSize=64, miniBatchSize=32, epochs=1)
    ppo = PT.TorchPPOLearner(11, 90, cfg, "cpu")
    import copy
    pol0, val0 = copy.deepcopy(ppo.policy), copy.deepcopy(ppo.value_net)
    exp = PT.ExperienceBuffer(64, 0, "cpu")
    rows = _fake_rows(64, 11, 3)
    exp.submit(rows)
    rep = {}
    ppo.learn(exp, rep)
    assert rep["Cumulative Model Updates"] == 1 and np.isfinite(rep["Policy Entropy"]) and rep["Policy Update Magnitude"] > 0
    # independent recomputation: gradient accumulation over the two minibatches of the SAME shuffled batch, clip 0.5, Adam
    exp2 = PT.ExperienceBuffer(64, 0, "cpu")
    exp2.submit(rows)
    batch = next(exp2.get_all_batches_shuffled(64))
    op, ov = torch.optim.Adam(pol0.parameters(), lr=cfg.policyLR), torch.optim.Adam(val0.parameters(), lr=cfg.criticLR)
    for s in (0, 32):
        mb = {k: v[s:s + 32] for k, v in batch.items()}
        a, b = _restated_losses(pol0, val0, mb, cfg, 0.5)
        a.backward(); b.backward()
    torch.nn.utils.clip_grad_norm_(pol0.parameters(), 0.5); torch.nn.utils.clip_grad_norm_(val0.parameters(), 0.5)
    op.step(); ov.step()
    for p, q in zip(ppo.policy.parameters(), pol0.parameters()):
        assert torch.allclose(p, q, atol=1e-6), (p - q).abs().max()
    for p, q in zip(ppo.value_net.parameters(), val0.parameters()):
        assert torch.allclose(p, q, atol=1e-6)


def test_update_learning_rates():
    """PPOLearner::UpdateLearningRates (PPOLearner.cpp:504-517): config and both Adam groups; a zero rate freezes that network
    (Learn skips its optimiser step, PPOLearner.cpp:262-281)."""
    torch.manual_seed(1)
    cfg = L.PPOLearnerConfig(policyLayerSizes=[16], criticLayerSizes=[16], batchSize=32, miniBatchSize=32, epochs=1)
    ppo = PT.TorchPPOLearner(5, 90, cfg, "cpu")
    ppo.update_learning_rates(0.0, 3e-3)
    assert cfg.policyLR == 0.0 and cfg.criticLR == 3e-3
    assert all(g["lr"] == 0.0 for g in ppo.policy_opt.param_groups) and all(g["lr"] == 3e-3 for g in ppo.value_opt.param_groups)
    pol0 = [p.detach().clone() for p in ppo.policy.parameters()]
    val0 = [p.detach().clone() for p in ppo.value_net.parameters()]
    exp = PT.ExperienceBuffer(32, 0, "cpu")
    exp.submit(_fake_rows(32, 5, 4))
    ppo.learn(exp, {})
    assert all(torch.equal(a, b) for a, b in zip(pol0, ppo.policy.parameters()))
    assert any(not torch.equal(a, b) for a, b in zip(val0, ppo.value_net.parameters()))


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)  # different init per rank on purpose: the learner must broadcast rank 0's
    cfg = L.PPOLearnerConfig(policyLayerSizes=[32], criticLayerSizes=[32], batchSize=32, miniBatchSize=32, epochs=2)
    ppo = PT.TorchPPOLearner(7, 90, cfg, "cpu")
    exp = PT.ExperienceBuffer(32, 5, "cpu")  # same shuffle seed on both ranks; different data shards
    exp.submit(_fake_rows(32, 7, 10 + rank))
    ppo.learn(exp, {})
    flat = torch.cat([p.detach().reshape(-1) for p in list(ppo.policy.parameters()) + list(ppo.value_net.parameters())])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save(gathered, out)
    dist.destroy_process_group()


def test_data_parallel_world2_gloo(tmp_path):
    out = str(tmp_path / "params.pt")
    mp.spawn(_dp_worker, args=(2, 29611, out), nprocs=2, join=True)
    g = torch.load(out)
    assert torch.equal(g[0], g[1]), "replicas diverged: gradients were not all-reduced identically"
    # and it equals a single process that sees BOTH shards per step with the averaged gradient
    torch.manual_seed(100)
    cfg = L.PPOLearnerConfig(policyLayerSizes=[32], criticLayerSizes=[32], batchSize=32, miniBatchSize=32, epochs=2)
    ref = PT.TorchPPOLearner(7, 90, cfg, "cpu")
    shards = [_fake_rows(32, 7, 10), _fake_rows(32, 7, 11)]
    exps = []
    for s in shards:
        e = PT.ExperienceBuffer(32, 5, "cpu"); e.submit(s); exps.append(e)
    for _ in range(cfg.epochs):
        batches = [next(e.get_all_batches_shuffled(32)) for e in exps]
        ref.policy_opt.zero_grad(); ref.value_opt.zero_grad()
        for b in batches:
            a, v = _restated_losses(ref.policy, ref.value_net, b, cfg, 1.0)
            (a / 2).backward(); (v / 2).backward()
        torch.nn.utils.clip_grad_norm_(ref.policy.parameters(), 0.5); torch.nn.utils.clip_grad_norm_(ref.value_net.parameters(), 0.5)
        ref.policy_opt.step(); ref.value_opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in list(ref.policy.parameters()) + list(ref.value_net.parameters())])
    assert torch.allclose(flat, g[0], atol=2e-6), (flat - g[0]).abs().max()


def test_shard_sizes_split_the_global_batch():
    """Data-parallel replicas: the reference's sizes are GLOBAL, every replica takes 1 / world of each (per-rank batch
    sizes must shrink with the buffer or a replica never forms a batch)."""
    cfg = L.LearnerConfig(expBufferSize=100_000, timestepsPerIteration=50_000)
    cfg.ppo.batchSize, cfg.ppo.miniBatchSize = 50_000, 25_000
    assert L.shard_sizes(cfg, 1) == {"batchSize": 50_000, "miniBatchSize": 25_000, "expBufferSize": 100_000}
    assert L.shard_sizes(cfg, 8) == {"batchSize": 6_250, "miniBatchSize": 3_125, "expBufferSize": 12_500}
    cfg.ppo.miniBatchSize = 0
    assert L.shard_sizes(cfg, 2)["miniBatchSize"] == 25_000
    with pytest.raises(RuntimeError, match="multiple"):
        L.shard_sizes(cfg, 3)
    cfg.expBufferSize = 40_000
    with pytest.raises(RuntimeError, match="smaller than"):
        L.shard_sizes(cfg, 1)


def test_device_learner_has_no_cpu_path():
    with pytest.raises(RuntimeError, match="CUDA"):
        L.PPOLearner(11, 90, L.PPOLearnerConfig(batchSize=64), "cpu")
