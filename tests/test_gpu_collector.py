"""-m gpu: the device-resident collector (policy/critic MLP on tcgen05, sampling, trajectory ring, GAE, row export)
through the C ABI, against oracle/ppo_oracle.py.

Tolerances (floating point, stated here as the task requires):
 * MLP forward: TF32 inputs (10-bit mantissa, round-to-nearest) with FP32 accumulation against an FP32 numpy forward:
   |value - ref| <= 2e-2 + 1e-2*|ref|, |logprob - ref_logprob[action]| <= 3e-2. (The reference runs libtorch FP32; its
   GEMM summation order is unspecified, so there is no bit-exact target.)
 * GAE / returns / value targets: BIT-EXACT against the scalar restatement of TorchFuncs::ComputeGAE.
 * trajectory ring vs. the plain step API: bit-exact."""
import numpy as np
import pytest

from oracle import ppo_oracle as po
from rlgymppo_cpp_b200 import abi, collector, engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available()
    return torch


def _mk(n_arenas, team=1, max_steps=4, seed=123, **kw):
    cfg = abi.default_cfg(num_arenas=n_arenas, team_size=team)
    e = engine.Engine(cfg)
    c = collector.Collector(e, max_steps=max_steps, seed=seed, **kw)
    c.init_default(seed=7)
    return e, c


def _infer(torch, c, obs_np, counter=0):
    n = obs_np.shape[0]
    obs = torch.from_numpy(obs_np).cuda()
    act = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    lp = torch.full((n,), 7.0, dtype=torch.float32, device="cuda")
    val = torch.full((n,), 7.0, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    c.infer(obs.data_ptr(), n, counter, act.data_ptr(), lp.data_ptr(), val.data_ptr())
    c.engine.sync()
    return act.cpu().numpy(), lp.cpu().numpy(), val.cpu().numpy()


@pytest.mark.parametrize("n_rows", [1, 127, 128, 1000])
def test_mlp_forward_matches_fp32_reference(torch_cuda, n_rows):
    e, c = _mk(64)
    rng = np.random.default_rng(n_rows)
    obs = rng.uniform(-1.5, 1.5, size=(n_rows, e.obs_size)).astype(np.float32)
    act, lp, val = _infer(torch_cuda, c, obs)
    ref_val = po.mlp_forward(c.weights[1], obs)[:, 0]
    ref_p = po.policy_probs(po.mlp_forward(c.weights[0], obs))
    assert np.all((act >= 0) & (act < 90))
    assert np.all(np.abs(val - ref_val) <= 2e-2 + 1e-2 * np.abs(ref_val)), np.abs(val - ref_val).max()
    ref_lp = np.log(ref_p[np.arange(n_rows), act])
    assert np.all(np.abs(lp - ref_lp) <= 3e-2), np.abs(lp - ref_lp).max()


def test_mlp_small_hidden_and_scaled_weights(torch_cuda):
    """non-default layer widths (64, 128) and sharper logits."""
    cfg = abi.default_cfg(num_arenas=8, team_size=2)
    e = engine.Engine(cfg)
    c = collector.Collector(e, policy_hidden=(64, 128), critic_hidden=(128, 64), max_steps=2, temperature=0.5)
    pol = collector.default_linear_init(c.policy_dims, 3)
    pol = [(W * 2.0, b) for W, b in pol]
    c.set_weights(0, pol)
    c.set_weights(1, collector.default_linear_init(c.critic_dims, 4))
    obs = np.random.default_rng(0).uniform(-1, 1, size=(300, e.obs_size)).astype(np.float32)
    act, lp, val = _infer(torch_cuda, c, obs)
    ref_val = po.mlp_forward(c.weights[1], obs)[:, 0]
    ref_p = po.policy_probs(po.mlp_forward(c.weights[0], obs), temperature=0.5)
    assert np.all(np.abs(val - ref_val) <= 2e-2 + 1e-2 * np.abs(ref_val))
    assert np.all(np.abs(lp - np.log(ref_p[np.arange(300), act])) <= 3e-2)


def test_deterministic_is_argmax(torch_cuda):
    e, c = _mk(64, deterministic=True)
    obs = np.random.default_rng(5).uniform(-1.5, 1.5, size=(512, e.obs_size)).astype(np.float32)
    act, lp, _ = _infer(torch_cuda, c, obs)
    ref_p = po.policy_probs(po.mlp_forward(c.weights[0], obs))
    top2 = np.sort(ref_p, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 2e-3 * top2[:, 1]  # rows whose argmax is not a near-tie at TF32 precision
    assert clear.sum() > 100
    assert np.array_equal(act[clear], ref_p.argmax(axis=1)[clear])
    assert np.all(lp == 0)  # DiscretePolicy.cpp:49-51


def test_sampling_distribution_and_streams(torch_cuda):
    e, c = _mk(64)
    row = np.random.default_rng(9).uniform(-1.5, 1.5, size=(1, e.obs_size)).astype(np.float32)
    n = 1 << 16
    obs = np.repeat(row, n, axis=0)
    a0, lp0, _ = _infer(torch_cuda, c, obs, counter=11)
    a0b, _, _ = _infer(torch_cuda, c, obs, counter=11)
    a1, _, _ = _infer(torch_cuda, c, obs, counter=12)
    assert np.array_equal(a0, a0b)          # counter-based: same (seed, counter, row) -> same sample
    assert (a0 != a1).mean() > 0.5           # a new step counter is a new stream
    p = po.policy_probs(po.mlp_forward(c.weights[0], row))[0]
    emp = np.bincount(a0, minlength=90) / n
    assert 0.5 * np.abs(emp - p).sum() < 0.02, 0.5 * np.abs(emp - p).sum()  # total variation distance


def test_collect_ring_consistent_with_step_api_and_gae_bit_exact(torch_cuda):
    torch = torch_cuda
    A, T = 256, 5
    e1, c1 = _mk(A, max_steps=T)
    e1.reset()
    c1.collect(T)
    c1.gae(0.99, 0.95, 2.0, 10.0)
    e1.sync()
    obs, act, rew, done, val = (c1.read(k) for k in ("obs", "action", "reward", "done", "value"))
    lp = c1.read("logprob")
    assert np.isfinite(obs).all() and np.isfinite(val).all() and np.isfinite(lp).all() and (lp <= 0).all()
    # (1) same engine seed stepped through the plain API with the ring's actions reproduces the ring bit for bit
    e2 = engine.Engine(abi.default_cfg(num_arenas=A, team_size=1))
    e2.reset()
    o0, _, _ = e2.read_outputs()
    assert np.array_equal(o0, obs[0])
    for t in range(T):
        a = torch.from_numpy(act[t]).cuda()
        e2.step_device(a.data_ptr())
        o, r, d = e2.read_outputs()
        assert np.array_equal(o, obs[t + 1]) and np.array_equal(r, rew[t]) and np.array_equal(d, done[t])
    # (2) values/logprobs in the ring are what a stand-alone inference of the same obs gives
    ref_val = po.mlp_forward(c1.weights[1], obs.reshape(-1, e1.obs_size))[:, 0].reshape(T + 1, -1)
    assert np.all(np.abs(val - ref_val) <= 2e-2 + 1e-2 * np.abs(ref_val))
    # (3) GAE bit-exact vs the scalar restatement over the reference's concatenation order (incl. the seam quirk)
    adv, tgt, ret = (c1.read(k) for k in ("advantage", "value_target", "ret"))
    o_adv, o_tgt, o_ret = po.gae_reference_order(rew, done, val, e1.P, 0.99, 0.95, 2.0, 10.0)
    assert np.array_equal(adv.view(np.uint32), o_adv.view(np.uint32))
    assert np.array_equal(tgt.view(np.uint32), o_tgt.view(np.uint32))
    assert np.array_equal(ret.view(np.uint32), o_ret.view(np.uint32))
    # (4) export in ExperienceBuffer row order
    N = A * e1.P
    st = torch.empty((N * T, e1.obs_size), dtype=torch.float32, device="cuda")
    nx = torch.empty_like(st)
    ac = torch.empty(N * T, dtype=torch.int64, device="cuda")
    f = lambda: torch.empty(N * T, dtype=torch.float32, device="cuda")
    lpx, rw, dn, tr, vt, ad = f(), f(), f(), f(), f(), f()
    c1.export_rows(st.data_ptr(), ac.data_ptr(), lpx.data_ptr(), rw.data_ptr(), nx.data_ptr(), dn.data_ptr(), tr.data_ptr(), vt.data_ptr(), ad.data_ptr())
    e1.sync()
    cat = po.concat_reference_order({"states": obs[:T], "next": obs[1:], "act": act, "lp": lp, "rw": rew, "vt": tgt, "ad": adv}, done, e1.P)
    assert np.array_equal(st.cpu().numpy(), cat["states"]) and np.array_equal(nx.cpu().numpy(), cat["next"])
    assert np.array_equal(ac.cpu().numpy(), cat["act"].astype(np.int64))
    assert np.array_equal(lpx.cpu().numpy(), cat["lp"]) and np.array_equal(rw.cpu().numpy(), cat["rw"])
    assert np.array_equal(dn.cpu().numpy(), cat["dones"]) and np.array_equal(tr.cpu().numpy(), cat["truncateds"])
    assert np.array_equal(vt.cpu().numpy(), cat["vt"]) and np.array_equal(ad.cpu().numpy(), cat["ad"])
    # (5) a second collect continues from the last observation
    c1.collect(2)
    e1.sync()
    assert np.array_equal(c1.read("obs")[0], obs[T])


def test_inference_under_the_tail_of_the_step_changes_nothing(torch_cuda, monkeypatch):
    """Full-size pool (one role block per SM, 256 inference tiles): the collect loop with every inference launched as the programmatic
    dependent of the step before it (per-block ready flags, rlg_engine_step_ready) fills the ring with exactly what the serialised
    launches produce — observations, sampled actions, log-probs, values, rewards, done flags."""
    rings = []
    for overlap, chain in (("0", "0"), ("1", "0"), ("1", "1")):  # serialised | inference under the step's tail | + per-block chains (opt-in)
        monkeypatch.setenv("RLG_COLLECT_OVERLAP", overlap)
        monkeypatch.setenv("RLG_COLLECT_CHAIN", chain)
        e, c = _mk(16384, max_steps=3)
        e.reset()
        c.collect(3)
        c.collect(3)  # the second collect starts from the ring's last slot
        e.sync()
        rings.append({k: c.read(k).copy() for k in ("obs", "action", "logprob", "value", "reward", "done")})
        flags, seq, apb = e.step_ready()
        assert seq == 6 and apb >= 32 and flags != 0
        del c, e
    for other in rings[1:]:
        for k in rings[0]:
            assert np.array_equal(rings[0][k].view(np.uint8), other[k].view(np.uint8)), k


def test_user_action_table_replaces_discrete_action(torch_cuda):
    """A user ActionParser as its table (rlg_engine_set_action_table): (1) a permutation of DiscreteAction's rows with the indices mapped
    through it reproduces the default engine bit for bit; (2) a narrower table narrows the policy head (GetActionAmount) and the collect
    loop only samples its indices; (3) bad tables are refused."""
    torch = torch_cuda
    A = 96
    rng = np.random.default_rng(5)
    perm = rng.permutation(90)
    e0 = engine.Engine(abi.default_cfg(num_arenas=A, team_size=1))
    e1 = engine.Engine(abi.default_cfg(num_arenas=A, team_size=1))
    table = engine.action_table()
    e1.set_action_table(table[perm])          # row j of the user table = DiscreteAction row perm[j]
    inv = np.argsort(perm)                    # DiscreteAction index i sits at user index inv[i]
    assert e1.num_actions == 90 and np.array_equal(e1.action_table, table[perm])
    e0.reset(); e1.reset()
    for s in range(12):
        a = rng.integers(0, 90, size=A * 2).astype(np.int32)
        o0, r0, d0 = e0.step_host(a)
        o1, r1, d1 = e1.step_host(inv[a].astype(np.int32))
        assert np.array_equal(o0.view(np.uint32), o1.view(np.uint32)) and np.array_equal(r0.view(np.uint32), r1.view(np.uint32)) and np.array_equal(d0, d1)
    e2 = engine.Engine(abi.default_cfg(num_arenas=A, team_size=1))
    e2.set_action_table(table[:37])
    c2 = collector.Collector(e2, max_steps=4, seed=3)
    c2.init_default(seed=7)
    assert c2.policy_dims[-1][0] == 37
    e2.reset()
    c2.collect(4)
    e2.sync()
    act = c2.read("action")
    assert act.min() >= 0 and act.max() < 37 and len(np.unique(act)) > 20
    with pytest.raises(engine.EngineError):
        e2.set_action_table(np.zeros((97, 8), np.float32))
    with pytest.raises(engine.EngineError):
        e2.set_action_table(np.full((4, 8), np.nan, np.float32))


def test_collector_argument_errors(torch_cuda):
    e = engine.Engine(abi.default_cfg(num_arenas=4, team_size=1))
    with pytest.raises(engine.EngineError):
        collector.Collector(e, policy_hidden=(100,), critic_hidden=(128,))  # not a multiple of 32
    c = collector.Collector(e, max_steps=2)
    with pytest.raises(engine.EngineError):
        c.collect(1)  # weights never set
    c.init_default()
    with pytest.raises(engine.EngineError):
        c.collect(3)  # > max_steps
    with pytest.raises(engine.EngineError):
        c.gae()  # nothing collected yet


def test_inference_kernel_against_the_reference_binary_fixture(torch_cuda, golden_dir):
    """k_mlp_infer on the network / observations of tests/golden/ppo_reference.npz — the outputs of the reference's own
    DiscretePolicy::GetAction(deterministic) and ValueEstimator::Forward (libtorch fp32): same argmax wherever the top two
    probabilities are further apart than the TF32 error, values within 2e-3."""
    import os

    g = np.load(os.path.join(golden_dir, "ppo_reference.npz"))
    e = engine.Engine(abi.default_cfg(num_arenas=32, team_size=1))
    c = collector.Collector(e, policy_hidden=(64, 64), critic_hidden=(64, 64), max_steps=1, seed=1, deterministic=True)
    layers = [(g[f"net/W{l}"], g[f"net/b{l}"]) for l in range(3)]
    c.set_weights(0, layers)
    c.set_weights(1, layers[:-1] + [(g["net/Wc"], g["net/bc"])])
    act, lp, val = _infer(torch_cuda, c, np.ascontiguousarray(g["net/obs"]))
    ref = g["policy_t1/probs"]
    top2 = np.sort(ref, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 2e-3
    assert clear.sum() > 20 and np.array_equal(act[clear], g["policy_t1/argmax"][clear])
    assert np.abs(val - g["critic/values"]).max() < 2e-3 * max(1.0, float(np.abs(g["critic/values"]).max()))
    assert np.all(lp == 0)  # deterministic: GetAction returns zeros for the log-probs (DiscretePolicy.cpp:49-52)


def test_python_host_state_setter(torch_cuda):
    """A user StateSetter in Python (BASELINE configs[3]: "custom StateSetter"): every episode starts from the state it
    writes — at GameInst::Start and whenever an arena's episode ends during collection."""
    cfg = abi.default_cfg(num_arenas=64, team_size=1)
    cfg.state_setter = abi.RLG_SETTER_HOST
    cfg.no_touch_max_steps = 3  # episodes end every 3 steps
    e = engine.Engine(cfg)
    c = collector.Collector(e, policy_hidden=(64, 64), critic_hidden=(64, 64), max_steps=8, seed=3)
    c.init_default(seed=3)
    calls = []

    def setter(ids, cars, balls):
        calls.append(len(ids))
        balls["pos"][:] = (0.0, 0.0, 500.0)
        balls["vel"][:] = (0.0, 0.0, 0.0)
        cars["pos"][:, 0] = (-1000.0, 0.0, 17.0)
        cars["pos"][:, 1] = (1000.0, 0.0, 17.0)
        cars["boost"][:] = 42.0

    c.set_state_setter(setter)
    c.reset_with_setter()
    assert calls == [64]
    cars, balls, pads, _ = e.get_state(np.arange(64, dtype=np.int32))
    assert np.allclose(balls["pos"], (0, 0, 500)) and np.allclose(cars["boost"], 42.0) and np.all(pads["is_active"] == 1)
    c.collect(8)
    e.sync()
    done = c.read("done")  # [T, A]
    assert done[2].all() and done[5].all() and not done[0].any()  # NoTouchCondition(3): every arena ends at steps 3 and 6
    assert calls[1:] == [64, 64]
    # the observation that follows a finished step is the custom reset state: ball at (0, 0, 500) / (4096, 5120, 2044)
    obs = c.read("obs")  # [T+1, N, obs]
    assert np.allclose(obs[3][:, 2], 500.0 / 2044.0, atol=1e-6) and np.allclose(obs[6][:, 0:2], 0.0, atol=1e-7)
