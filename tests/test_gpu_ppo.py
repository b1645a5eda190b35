"""-m gpu: the device PPO learner (csrc/ppo.cu through rlgymppo_cpp_b200.ppo / learner.PPOLearner) against the torch fp32
restatement of PPOLearner::Learn (oracle/ppo_torch.py, CPU) from identical weights, buffer contents and batch permutations.

Tolerances.  The device runs every GEMM with TF32 inputs / FP32 accumulation (SURVEY 8d allows it).  Against a PLAIN fp32
restatement the difference is not small element-wise: a hidden unit whose pre-activation lies within TF32 rounding of zero takes
its ReLU-backward decision the other way and switches a whole row's term of that unit's weight gradient (measured: ~2-4 % of a
tensor's largest gradient, at any batch size), and Adam's first step turns a flipped gradient sign into 2 * lr.  So the sharp
comparisons use the restatement with its contractions' operands truncated to TF32 (oracle/ppo_torch.py TF32Linear: the same
arithmetic up to fp32 summation order): gradients within 3e-4 of each tensor's largest entry, parameter updates with cosine
similarity > 0.999 and < 0.2 % of the parameters further than 15 % of a step apart.  One plain-fp32 comparison stays, with the
bounds that the effect above allows (cosine > 0.98, < 1 % of the parameters off); reported diagnostics within 2e-3 relative.  Adam and clip_grad_norm_ are torch's own operators in the restatement (unpinned against the
reference binary by construction: the reference calls the same libtorch code)."""
import os

import numpy as np
import pytest

from oracle import ppo_oracle as po
from oracle import ppo_torch as PT
from rlgymppo_cpp_b200 import learner as L

pytestmark = pytest.mark.gpu


def _rows(n, obs, seed, policy=None):
    """Synthetic experience; the old log-probs come from the policy itself (ratio ~ 1, some rows pushed over the clip range)."""
    import torch

    g = np.random.default_rng(seed)
    states = g.normal(size=(n, obs)).astype(np.float32)
    actions = g.integers(0, 90, size=n)
    if policy is not None:
        with torch.no_grad():
            lp = torch.log_softmax(policy(torch.from_numpy(states)), -1).numpy()[np.arange(n), actions]
        # ratio = exp(lp - old): spread over both sides of the clip range, but no row within 1 % of a clip boundary — the clipped
        # surrogate's gradient is discontinuous there and a TF32-sized difference in one logit would switch a whole row on or off
        delta = g.normal(scale=0.15, size=n)
        for edge in (np.log(0.8), np.log(1.2)):
            near = np.abs(-delta - edge) < 0.01
            delta[near] += 0.03 * np.sign(-delta[near] - edge + 1e-12) * -1
        log_probs = (lp + delta).astype(np.float32)
    else:
        log_probs = np.log(g.uniform(0.005, 0.05, size=n)).astype(np.float32)
    return {"states": states, "actions": actions.astype(np.int64), "log_probs": log_probs, "values": g.normal(size=n).astype(np.float32),
            "advantages": g.normal(size=n).astype(np.float32)}


def ours_stream_sync(dev):
    """rlg_ppo_submit ran on the learner's own stream: wait for it before the source tensors go away."""
    import torch

    torch.cuda.synchronize()


def _flat(ppo):
    import torch

    return torch.cat([p.detach().reshape(-1).cpu() for p in list(ppo.policy.parameters()) + list(ppo.value_net.parameters())]).clone()


def _compare_updates(init, ours, ref, lr, frac_bad=0.01, cos_min=0.995, close=0.15):
    import torch

    du, dr = ours - init, ref - init
    cos = float(torch.dot(du, dr) / (du.norm() * dr.norm() + 1e-30))
    bad = float(((ours - ref).abs() > close * lr).float().mean())
    print(f"update cosine {cos:.6f}, fraction of parameters further than {close} steps apart {bad:.5f}, max diff {float((ours - ref).abs().max()):.3e}")
    assert cos > cos_min and bad < frac_bad, (cos, bad)


def _submit(dev, rows):
    import torch

    t = {k: torch.from_numpy(v).cuda() for k, v in rows.items()}
    torch.cuda.synchronize()
    dev.submit(t["states"].data_ptr(), t["actions"].data_ptr(), t["log_probs"].data_ptr(), t["values"].data_ptr(), t["advantages"].data_ptr(), len(rows["actions"]))
    ours_stream_sync(dev)


def _make_pair(obs, hidden, batch, mini, epochs, seed, ent=0.01, lr=2e-4, exp_size=None, tf32=True):
    import torch

    torch.manual_seed(seed)
    cfg = L.PPOLearnerConfig(policyLayerSizes=list(hidden), criticLayerSizes=list(hidden), batchSize=batch, miniBatchSize=mini, epochs=epochs,
                             policyLR=lr, criticLR=lr, entCoef=ent)
    ours = L.PPOLearner(obs, 90, cfg, "cuda:0", exp_buffer_size=exp_size or batch, seed=seed)
    import copy

    ref = PT.TorchPPOLearner(obs, 90, copy.deepcopy(cfg), "cpu", emulate_tf32=tf32)
    ref.policy.load_state_dict(ours.policy.state_dict())
    ref.value_net.load_state_dict(ours.value_net.state_dict())
    return ours, ref, cfg


def test_device_experience_buffer_is_the_reference_fifo():
    """ExperienceBuffer.cpp:12-70 on the device ring: same logical contents as the numpy oracle (pinned to the reference binary in
    test_ppo_oracle.py) after fills, overflows and an oversize submit."""
    import torch

    torch.manual_seed(0)
    cfg = L.PPOLearnerConfig(policyLayerSizes=[32], criticLayerSizes=[32], batchSize=4, miniBatchSize=4)
    ours = L.PPOLearner(7, 90, cfg, "cuda:0", exp_buffer_size=10, seed=1)
    o = po.ExperienceBufferOracle(10)
    for i, n in enumerate([4, 4, 5, 13, 2, 9]):
        rows = _rows(n, 7, i)
        _submit(ours.dev, rows)
        o.submit(rows)
        got = ours.dev.buffer_read()
        assert ours.dev.buffer_size == o.cur
        for k in ("states", "actions", "log_probs", "values", "advantages"):
            assert np.array_equal(got[k], o.data[k][: o.cur]), (i, k)
    perm = ours.dev.peek_shuffle()
    assert sorted(perm.tolist()) == list(range(10))
    assert not np.array_equal(perm, ours.dev.peek_shuffle(counter=5))


@pytest.mark.parametrize("hidden,obs,batch,mini,tf32", [((64, 64), 89, 512, 256, True), ((256, 256, 256), 89, 4096, 1024, True), ((32,), 70, 256, 256, True),
                                                        ((256, 256, 256), 89, 4096, 1024, False)])
def test_learn_step_matches_the_torch_restatement(hidden, obs, batch, mini, tf32):
    import torch

    ours, ref, cfg = _make_pair(obs, hidden, batch, mini, 1, seed=3, tf32=tf32)
    rows = _rows(batch, obs, 11, policy=ref.policy)
    _submit(ours.dev, rows)
    exp = PT.ExperienceBuffer(batch, 0, "cpu")
    exp.submit({k: torch.from_numpy(v) for k, v in rows.items()})
    exp.forced_perms = [ours.dev.peek_shuffle()]
    rep_o, rep_r = {}, {}
    init = _flat(ref)
    ours.learn(rep_o)
    ref.learn(exp, rep_r)
    assert rep_o["Cumulative Model Updates"] == rep_r["Cumulative Model Updates"] == 1
    print({k: (rep_o[k], rep_r[k]) for k in ("Policy Entropy", "Mean KL Divergence", "Mean Ratio", "Value Function Loss", "SB3 Clip Fraction")})
    if tf32:
        _compare_updates(init, _flat(ours), _flat(ref), 2e-4, frac_bad=0.002, cos_min=0.999)
    else:
        _compare_updates(init, _flat(ours), _flat(ref), 2e-4, frac_bad=0.01, cos_min=0.98)
    for k in ("Policy Entropy", "Mean Ratio", "Value Function Loss", "Policy Update Magnitude", "Value Function Update Magnitude"):
        assert abs(rep_o[k] - rep_r[k]) <= 2e-3 * abs(rep_r[k]) + 1e-6, (k, rep_o[k], rep_r[k])
    assert abs(rep_o["Mean KL Divergence"] - rep_r["Mean KL Divergence"]) < 2e-4
    assert abs(rep_o["SB3 Clip Fraction"] - rep_r["SB3 Clip Fraction"]) < 5e-3
    assert 0.05 < rep_r["SB3 Clip Fraction"] < 0.95  # the clip branch is exercised on both sides


def test_gradients_match_autograd():
    """The hand-written backward (loss kernels + GEMMs + bias sums, accumulated over two minibatches) against torch autograd on
    the TF32-operand restatement's losses: every gradient within 3e-4 of the largest gradient of its tensor."""
    import torch

    from rlgymppo_cpp_b200 import ppo as P

    ours, ref, cfg = _make_pair(89, (128, 64), 512, 256, 1, seed=5, lr=0.0)  # lr 0 on both nets would skip; use the raw pieces instead
    ours.update_learning_rates(1e-30, 1e-30)  # train (gradients flow), parameters do not move
    ref.update_learning_rates(1e-30, 1e-30)
    rows = _rows(512, 89, 12, policy=ref.policy)
    _submit(ours.dev, rows)
    perm = ours.dev.peek_shuffle()
    # autograd on the restatement
    t = {k: torch.from_numpy(v[perm]) for k, v in rows.items()}
    acc = torch.zeros(5)
    for s in (0, 256):
        ref._minibatch(t["states"][s:s + 256], t["actions"][s:s + 256], t["advantages"][s:s + 256], t["log_probs"][s:s + 256], t["values"][s:s + 256], acc)
    # device: intercept the gradients through the all-reduce hook (called once per batch, before the optimiser step)
    grabbed = {}

    def hook(ptr, count, stream):
        torch.cuda.synchronize()
        grabbed["pol"] = ours.dev.get_layers(0, P.GRADS)
        grabbed["val"] = ours.dev.get_layers(1, P.GRADS)

    ours.dev.set_allreduce(hook, 2)  # world 2 so that the hook runs; the flat gradient itself is left untouched
    ours.learn({})
    assert grabbed
    for name, seq in (("pol", ref.policy), ("val", ref.value_net)):
        lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
        for l, (m, (gw, gb)) in enumerate(zip(lin, grabbed[name])):
            for got, want in ((gw, m.weight.grad.numpy()), (gb, m.bias.grad.numpy())):
                scale = float(np.abs(want).max()) + 1e-12
                assert float(np.abs(got - want).max()) <= 3e-4 * scale, (name, l, float(np.abs(got - want).max()), scale)


def test_many_steps_track_the_restatement_and_frozen_networks_stay_put():
    """10 optimiser steps over 2 epochs with buffer overflow in between: the two learners stay within TF32 drift (1e-3); with
    policyLR = 0 the policy is bit-identical after Learn while the critic moves (PPOLearner.cpp:262-281), and un-freezing works."""
    import torch

    ours, ref, cfg = _make_pair(89, (64, 64), 256, 128, 2, seed=9, exp_size=1280)
    exp = PT.ExperienceBuffer(1280, 0, "cpu")
    init = _flat(ref)
    for it in range(2):
        rows = _rows(1024, 89, 20 + it, policy=ref.policy)
        _submit(ours.dev, rows)
        exp.submit({k: torch.from_numpy(v) for k, v in rows.items()})
        c = ours.dev.L.rlg_ppo_shuffle_counter(ours.dev.h)
        exp.forced_perms = [ours.dev.peek_shuffle(counter=c + e) for e in range(2)]
        ro, rr = {}, {}
        ours.learn(ro)
        ref.learn(exp, rr)
        assert ro["Cumulative Model Updates"] == rr["Cumulative Model Updates"]
    assert ro["Cumulative Model Updates"] == 8 + 10
    _compare_updates(init, _flat(ours), _flat(ref), 18 * 2e-4, frac_bad=0.01, cos_min=0.998, close=0.1)  # 18 steps: within 10 % of the total path
    for k in ("Policy Entropy", "Value Function Loss", "Mean Ratio"):
        assert abs(ro[k] - rr[k]) <= 5e-3 * abs(rr[k]) + 1e-5, (k, ro[k], rr[k])
    p0 = [p.detach().clone() for p in ours.policy.parameters()]
    v0 = [p.detach().clone() for p in ours.value_net.parameters()]
    ours.update_learning_rates(0.0, 3e-4)
    ours.learn({})
    assert all(torch.equal(a, b) for a, b in zip(p0, ours.policy.parameters()))
    assert any(not torch.equal(a, b) for a, b in zip(v0, ours.value_net.parameters()))
    ours.update_learning_rates(3e-4, 3e-4)  # un-freeze: the very next Learn trains the policy again (stale-graph case)
    rep = {}
    ours.learn(rep)
    assert any(not torch.equal(a, b) for a, b in zip(p0, ours.policy.parameters())) and rep["SB3 Clip Fraction"] >= 0


def test_too_few_rows_is_loud_and_takes_no_step(capsys):
    import torch

    ours, ref, cfg = _make_pair(89, (32,), 256, 256, 1, seed=2, exp_size=512)
    _submit(ours.dev, _rows(100, 89, 1))
    p0 = [p.detach().clone() for p in ours.policy.parameters()]
    rep = {}
    ours.learn(rep)
    assert "no optimiser step" in capsys.readouterr().out
    assert all(torch.equal(a, b) for a, b in zip(p0, ours.policy.parameters())) and rep["Cumulative Model Updates"] == 0


def _nccl_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    torch.manual_seed(100 + rank)  # different init per rank on purpose: the learner must broadcast rank 0's
    cfg = L.PPOLearnerConfig(policyLayerSizes=[64, 64], criticLayerSizes=[64, 64], batchSize=256, miniBatchSize=128, epochs=2, policyLR=2e-4, criticLR=2e-4)
    ppo = L.PPOLearner(89, 90, cfg, f"cuda:{rank}", exp_buffer_size=256, seed=5)  # same shuffle seed on both ranks; different data shards
    init = torch.cat([p.detach().reshape(-1) for p in list(ppo.policy.parameters()) + list(ppo.value_net.parameters())])
    _submit_on(ppo.dev, _rows(256, 89, 10 + rank), rank)
    perms = [ppo.dev.peek_shuffle(counter=e) for e in range(2)]
    ppo.learn({})
    flat = torch.cat([p.detach().reshape(-1) for p in list(ppo.policy.parameters()) + list(ppo.value_net.parameters())]).cuda(rank)
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save({"params": [g.cpu() for g in gathered], "init": init, "perms": perms}, out)
    dist.destroy_process_group()


def _submit_on(dev, rows, rank):
    import torch

    t = {k: torch.from_numpy(v).cuda(rank) for k, v in rows.items()}
    torch.cuda.synchronize(rank)
    dev.submit(t["states"].data_ptr(), t["actions"].data_ptr(), t["log_probs"].data_ptr(), t["values"].data_ptr(), t["advantages"].data_ptr(), len(rows["actions"]))
    torch.cuda.synchronize(rank)


def test_data_parallel_two_ranks_nccl(tmp_path):
    """2 GPUs, NCCL: ONE all-reduce of the flat gradient per optimiser step through the hook; replicas end bit-identical and equal
    the restatement that sees both shards with the averaged gradient."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "params.pt")
    mp.spawn(_nccl_worker, args=(2, 29641, out), nprocs=2, join=True)
    g = torch.load(out, weights_only=False)
    assert torch.equal(g["params"][0], g["params"][1]), "replicas diverged"
    torch.manual_seed(100)
    cfg = L.PPOLearnerConfig(policyLayerSizes=[64, 64], criticLayerSizes=[64, 64], batchSize=256, miniBatchSize=128, epochs=2, policyLR=2e-4, criticLR=2e-4)
    ref = PT.TorchPPOLearner(89, 90, cfg, "cpu", emulate_tf32=True)
    flat0 = torch.cat([p.detach().reshape(-1) for p in list(ref.policy.parameters()) + list(ref.value_net.parameters())])
    assert torch.equal(flat0, g["init"]), "rank 0's initialisation was not the broadcast one"
    shards = [_rows(256, 89, 10), _rows(256, 89, 11)]
    for e in range(2):
        ref.policy_opt.zero_grad(); ref.value_opt.zero_grad()
        acc = torch.zeros(5)
        for rows in shards:
            t = {k: torch.from_numpy(v[g["perms"][e]]) for k, v in rows.items()}
            for s in (0, 128):
                ref._minibatch(t["states"][s:s + 128], t["actions"][s:s + 128], t["advantages"][s:s + 128], t["log_probs"][s:s + 128], t["values"][s:s + 128], acc)
        for p in list(ref.policy.parameters()) + list(ref.value_net.parameters()):
            p.grad /= 2
        torch.nn.utils.clip_grad_norm_(ref.policy.parameters(), 0.5); torch.nn.utils.clip_grad_norm_(ref.value_net.parameters(), 0.5)
        ref.policy_opt.step(); ref.value_opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in list(ref.policy.parameters()) + list(ref.value_net.parameters())])
    _compare_updates(g["init"], g["params"][0], flat, 2 * 2e-4, frac_bad=0.005, cos_min=0.998)
