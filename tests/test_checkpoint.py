"""Checkpoint layout of the reference (Learner.cpp:171-309, PPOLearner.cpp:362-502) written / read by rlgymppo_cpp_b200.checkpoint."""
import json
import os
import types

import numpy as np
import pytest
import torch

from rlgymppo_cpp_b200 import checkpoint as ck
from rlgymppo_cpp_b200 import learner as L


def _fake_learner(tmp_path, seed):
    torch.manual_seed(seed)
    cfg = L.LearnerConfig(checkpointSaveFolder=str(tmp_path), checkpointLoadFolder=str(tmp_path), checkpointsToKeep=2)
    ppo = types.SimpleNamespace(policy=L.make_mlp(89, [64, 32], 90), value_net=L.make_mlp(89, [64, 32], 1), cumulative_model_updates=7)
    ppo.policy_opt = torch.optim.Adam(ppo.policy.parameters(), lr=2e-4)
    ppo.value_opt = torch.optim.Adam(ppo.value_net.parameters(), lr=2e-4)
    ppo.policy(torch.zeros(4, 89)).sum().backward(); ppo.policy_opt.step()
    rs = L.WelfordRunningStat(); rs.increment(np.array([1.0, 2.0, 4.0], dtype=np.float32), 3)
    pushed = []
    return types.SimpleNamespace(cfg=cfg, ppo=ppo, total_timesteps=123456, total_epochs=9, return_stats=rs, skill_tracker=None,
                                 device=torch.device("cpu"), _push_weights=lambda: pushed.append(1), pushed=pushed)


def test_save_layout_and_round_trip(tmp_path):
    a = _fake_learner(tmp_path, 1)
    dst = ck.save_learner(a)
    assert os.path.basename(dst) == "123456"
    assert sorted(os.listdir(dst)) == sorted(["PPO_POLICY.lt", "PPO_CRITIC.lt", "PPO_POLICY_OPTIM.pt", "PPO_CRITIC_OPTIM.pt", "RUNNING_STATS.json"])
    j = json.load(open(os.path.join(dst, "RUNNING_STATS.json")))
    assert j["cumulative_timesteps"] == 123456 and j["cumulative_model_updates"] == 7 and j["epoch"] == 9
    assert j["reward_running_stats"]["shape"] == 1 and j["reward_running_stats"]["count"] == 3 and len(j["reward_running_stats"]["mean"]) == 1
    # the model file is a TorchScript archive with the parameter names torch::save(nn::Sequential) produces
    names = list(torch.jit.load(os.path.join(dst, "PPO_POLICY.lt")).state_dict().keys())
    assert names == ["0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias"]
    b = _fake_learner(tmp_path, 2)
    assert not torch.equal(next(a.ppo.policy.parameters()), next(b.ppo.policy.parameters()))
    assert ck.load_learner(b) == dst
    for x, y in zip(list(a.ppo.policy.parameters()) + list(a.ppo.value_net.parameters()), list(b.ppo.policy.parameters()) + list(b.ppo.value_net.parameters())):
        assert torch.equal(x, y)
    assert b.total_timesteps == 123456 and b.total_epochs == 9 and b.ppo.cumulative_model_updates == 7 and b.pushed == [1]
    assert b.return_stats.count == 3 and b.return_stats.get_std() == a.return_stats.get_std()
    sa, sb = a.ppo.policy_opt.state_dict()["state"], b.ppo.policy_opt.state_dict()["state"]
    assert all(torch.equal(sa[k]["exp_avg"], sb[k]["exp_avg"]) for k in sa)


def test_latest_is_loaded_and_old_ones_are_pruned(tmp_path):
    a = _fake_learner(tmp_path, 1)
    for ts in (100, 300, 200, 400):
        a.total_timesteps = ts
        ck.save_learner(a)
    assert ck.numbered_folders(str(tmp_path)) == [300, 400]  # checkpointsToKeep = 2: the lowest-numbered goes whenever there are more
    b = _fake_learner(tmp_path, 3)
    assert os.path.basename(ck.load_learner(b)) == "400" and b.total_timesteps == 400
    assert ck.load_learner(b, str(tmp_path / "nothing_here")) is None


def test_size_mismatch_is_an_error(tmp_path):
    a = _fake_learner(tmp_path, 1)
    dst = ck.save_learner(a)
    other = L.make_mlp(89, [64, 64], 90)
    with pytest.raises(RuntimeError, match="different size"):
        ck.load_seq(other, os.path.join(dst, "PPO_POLICY.lt"))
    with pytest.raises(RuntimeError, match="does not exist"):
        ck.load_seq(other, os.path.join(dst, "NOPE.lt"))


@pytest.fixture(scope="module")
def libtorch_tool(tmp_path_factory):
    """tests/ckpt_cpp/ckpt_roundtrip.cpp built against the pip libtorch: torch::save / torch::load of an nn::Sequential, i.e.
    the reference's TorchLoadSaveSeq (PPOLearner.cpp:372-419) on the reference's model layout."""
    import subprocess

    from torch.utils import cpp_extension as ce

    out = str(tmp_path_factory.mktemp("ckpt") / "ckpt_roundtrip")
    lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ckpt_cpp", "ckpt_roundtrip.cpp")
    cmd = (["g++", "-O1", "-std=c++17", src, "-o", out] + [f"-I{i}" for i in ce.include_paths()] +
           [f"-L{lib}", "-ltorch", "-ltorch_cpu", "-lc10", f"-Wl,-rpath,{lib}", "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI))])
    env = dict(os.environ); env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        pytest.skip("libtorch test tool does not build here: " + r.stderr[-300:])
    return out


def test_models_round_trip_through_libtorch(libtorch_tool, tmp_path):
    """Our PPO_POLICY.lt loads with libtorch's torch::load(nn::Sequential) (what the reference runs), and a file written by
    libtorch's torch::save(nn::Sequential) loads here — parameter by parameter."""
    import subprocess

    torch.manual_seed(11)
    seq = L.make_mlp(89, [64, 32], 90)
    ours = str(tmp_path / "PPO_POLICY.lt")
    ck.save_seq(seq, ours)
    r = subprocess.run([libtorch_tool, "load", ours, "89", "64,32", "90"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    rows = [l.split() for l in r.stdout.strip().splitlines()]
    params = list(seq.parameters())
    assert len(rows) == len(params)
    for (n, s1, s2), p in zip(rows, params):
        pd = p.detach().double()
        assert int(n) == p.numel()
        assert abs(float(s1) - float(pd.sum())) < 1e-5 and abs(float(s2) - float((pd * pd).sum())) < 1e-5
    theirs = str(tmp_path / "theirs.lt")
    r = subprocess.run([libtorch_tool, "save", theirs, "89", "64,32", "90"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    seq2 = L.make_mlp(89, [64, 32], 90)
    ck.load_seq(seq2, theirs)
    for i, p in enumerate(seq2.parameters()):
        assert torch.all(p == (i + 1) / 8)
