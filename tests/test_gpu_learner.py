"""-m gpu: the Learner loop over the device engine: collect -> GAE -> buffer -> PPO update -> weights back to the collector."""
import numpy as np
import pytest

from oracle import ppo_oracle as po
from rlgymppo_cpp_b200 import abi

pytestmark = pytest.mark.gpu


def test_learner_two_iterations_and_weight_roundtrip():
    import torch

    from rlgymppo_cpp_b200 import learner as L

    assert torch.cuda.is_available()
    cfg = L.LearnerConfig(numThreads=8, numGamesPerThread=32, timestepsPerIteration=2048, expBufferSize=6144, randomSeed=5,
                          ppo=L.PPOLearnerConfig(batchSize=2048, miniBatchSize=1024, epochs=2, policyLR=2e-4, criticLR=2e-4, entCoef=0.01))
    ecfg = abi.default_cfg(num_arenas=cfg.num_arenas, team_size=1)
    lr = L.Learner(ecfg, cfg)
    assert lr.steps_per_iter == 4  # 2048 / (256 arenas * 2 players)
    w0 = [p.detach().clone() for p in lr.ppo.policy.parameters()]
    reps = lr.learn(max_iterations=3)
    assert len(reps) == 3 and lr.total_timesteps == 3 * 2048
    for r in reps:
        for k in ("Policy Entropy", "Mean KL Divergence", "Value Function Loss", "Avg Return", "Collected Steps/Second"):
            assert np.isfinite(r[k]), (k, r[k])
        assert 0 < r["Policy Entropy"] <= np.log(90) + 1e-3
    assert abs(reps[0]["Mean Ratio"] - 1) < 5e-2  # first epoch starts at ratio ~1: TF32 inference vs the update's forward
    assert reps[1]["Cumulative Model Updates"] > reps[0]["Cumulative Model Updates"]
    assert any(not torch.equal(a, b) for a, b in zip(w0, lr.ppo.policy.parameters()))
    assert lr.ppo.dev.buffer_size == 6144
    # the collector now infers with the UPDATED weights
    obs = np.random.default_rng(0).uniform(-1, 1, size=(256, lr.engine.obs_size)).astype(np.float32)
    t_obs = torch.from_numpy(obs).cuda()
    val = torch.empty(256, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    lr.collector.infer(t_obs.data_ptr(), 256, 0, value_ptr=val.data_ptr())
    lr.engine.sync()
    ref = po.mlp_forward(L.mlp_layers_numpy(lr.ppo.value_net), obs)[:, 0]
    assert np.all(np.abs(val.cpu().numpy() - ref) <= 2e-2 + 1e-2 * np.abs(ref))


def test_skill_tracker_on_engine_detects_goals_and_rates():
    """SkillTracker over a device eval pool: a ball injected behind a goal line is a goal for the policy on the scoring side
    (Math::IsBallScored on the step's snapshot, SkillTracker.cpp:132-149) and ends the episode; RunGames moves the ELO."""
    import torch

    from rlgymppo_cpp_b200 import collector, skill_tracker as stm

    assert torch.cuda.is_available()
    stc = stm.SkillTrackerConfig(enabled=True, numEnvs=6, simTime=6 * 8 * 5 / 120, updateInterval=1, timestepsPerVersion=10 ** 9)
    ecfg = abi.default_cfg(num_arenas=64, team_size=1)
    st = stm.SkillTracker.on_engine(stc, ecfg, (64, 64), seed=7)
    e = st.engine
    assert e.A == 6 and st.mode == "1v1" and sorted(st.teams.tolist()) == [0, 1]
    # dummy reward: every reward is exactly 0
    done, scored = st.step_fn(np.zeros(e.A * e.P, dtype=np.int32))
    assert not scored.any()
    assert not e.read_outputs()[1].any()
    # ball behind the orange goal line (y > 0) in arena 1, behind the blue one in arena 4
    ids = np.array([1, 4], dtype=np.int32)
    cars, balls, pads, ticks = e.get_state(ids)
    balls["pos"][0] = (0.0, 5300.0, 100.0)
    balls["pos"][1] = (0.0, -5300.0, 100.0)
    balls["vel"][:] = 0
    e.set_state(ids, balls=balls)
    done, scored = st.step_fn(np.zeros(e.A * e.P, dtype=np.int32))
    assert scored.tolist() == [0, 1, 0, 0, -1, 0] and done[1] and done[4]
    # the finished arenas were re-set to kickoffs: no goal on the next step, ball back at the centre
    done, scored = st.step_fn(np.zeros(e.A * e.P, dtype=np.int32))
    assert not scored.any()
    _, b2, _, _ = e.get_state(ids)
    assert np.all(np.abs(b2["pos"][:, 1]) < 200)
    # RunGames with two different weight sets: plays 5 steps per game, freezes version 0, ratings stay finite and zero-sum
    w_cur = collector.default_linear_init(st.collector.policy_dims, 1)
    played = st.run_games(w_cur, 1000)
    assert played is not None and len(st.old_policies) == 1
    total = st.cur_rating["1v1"] + st.old_ratings[0]["1v1"]
    assert abs(total - 2000.0) < 1e-2


def test_learner_reports_skill_rating():
    import torch

    from rlgymppo_cpp_b200 import learner as L, skill_tracker as stm

    assert torch.cuda.is_available()
    stc = stm.SkillTrackerConfig(enabled=True, numEnvs=4, simTime=4 * 8 * 3 / 120, updateInterval=1, timestepsPerVersion=3000, maxVersions=2)
    cfg = L.LearnerConfig(numThreads=4, numGamesPerThread=32, timestepsPerIteration=1024, expBufferSize=2048, randomSeed=9, skillTrackerConfig=stc,
                          ppo=L.PPOLearnerConfig(batchSize=1024, miniBatchSize=512, epochs=1, policyLR=2e-4, criticLR=2e-4))
    lr = L.Learner(abi.default_cfg(num_arenas=cfg.num_arenas, team_size=1), cfg)
    reps = lr.learn(max_iterations=4)
    assert all("Skill Rating 1v1" in r and np.isfinite(r["Skill Rating 1v1"]) for r in reps)
    assert 1 <= len(lr.skill_tracker.old_policies) <= 2


def test_render_sender_document_from_engine():
    """RenderSender over a live engine: one arena's document has the reference's schema and the engine's values."""
    from rlgymppo_cpp_b200 import engine, sinks

    e = engine.Engine(abi.default_cfg(num_arenas=8, team_size=2))
    e.reset()
    acts = np.random.default_rng(0).integers(0, 90, size=e.A * e.P).astype(np.int32)
    e.step_host(acts)
    rs = sinks.RenderSender(e)
    doc = rs.document(arena=3, action_idx=acts[3 * e.P:4 * e.P])
    st = doc["state"]
    assert len(st["players"]) == 4 and len(st["boost_pads"]) == 34 and len(doc["actions"]) == 4 and len(doc["actions"][0]) == 8
    cars, balls, _, _ = e.get_state(np.array([3], dtype=np.int32))
    assert np.allclose(st["ball"]["pos"], balls[0]["pos"])
    assert sorted(p["team_num"] for p in st["players"]) == [0, 0, 1, 1]
    for p in st["players"]:
        assert np.allclose(p["phys"]["pos"], cars[0][p["car_id"] - 1]["pos"]) and 0 <= p["boost_amount"] <= 1
    rs.send(arena=3)  # UDP, nobody needs to listen


def test_infer_unit_policy_and_critic_from_a_checkpoint(tmp_path):
    """InferUnit (InferUnit.cpp:11-138) on the engine: a PPO_POLICY.lt / PPO_CRITIC.lt pair saved in the reference's format,
    loaded back, evaluated on the engine's current obs rows through the tcgen05 inference kernel, against torch fp32."""
    import torch

    from rlgymppo_cpp_b200 import checkpoint, engine, infer_unit
    from rlgymppo_cpp_b200.learner import make_mlp

    e = engine.Engine(abi.default_cfg(num_arenas=64, team_size=1))
    e.reset()
    e.step_host(np.random.default_rng(1).integers(0, 90, size=e.A * e.P).astype(np.int32))
    torch.manual_seed(5)
    sizes = [128, 128]
    pol, cri = make_mlp(e.obs_size, sizes, 90), make_mlp(e.obs_size, sizes, 1)
    checkpoint.save_seq(pol, str(tmp_path / "PPO_POLICY.lt"))
    checkpoint.save_seq(cri, str(tmp_path / "PPO_CRITIC.lt"))
    up = infer_unit.InferUnit(e, str(tmp_path / "PPO_POLICY.lt"), True, e.obs_size, sizes)
    uc = infer_unit.InferUnit(e, str(tmp_path / "PPO_CRITIC.lt"), False, e.obs_size, sizes)
    obs = up.get_obs()
    assert obs.shape == (e.A * e.P, e.obs_size)
    with torch.no_grad():
        logits = pol(torch.from_numpy(obs)).numpy()
        values = cri(torch.from_numpy(obs)).numpy().reshape(-1)
    # deterministic = argmax (ties / TF32 near-ties: the chosen logit is within 2e-3 of the best)
    idx = up.infer_policy_indices(True)
    assert idx.shape == (e.A * e.P,) and idx.min() >= 0 and idx.max() < 90
    assert np.all(logits.max(1) - logits[np.arange(len(idx)), idx] <= 2e-3 * np.maximum(1, np.abs(logits).max(1)))
    acts = up.infer_policy_all(True)
    assert acts.shape == (e.A * e.P, 8) and np.array_equal(acts, engine.action_table()[idx])
    # the same rows handed in by the caller, and one row alone
    assert np.array_equal(up.infer_policy_indices(True, obs=obs[:10]), idx[:10])
    assert np.array_equal(up.infer_policy_single(7, True), acts[7])
    # sampling: indices in range, not all equal to the argmax at a high temperature
    samp = up.infer_policy_indices(False, temperature=5.0)
    assert samp.min() >= 0 and samp.max() < 90 and (samp != idx).mean() > 0.5
    # distribution of one row
    p = up.infer_policy_single_distrib(3, temperature=2.0)
    ref = torch.softmax(torch.from_numpy(logits[3]) / 2.0, -1).clamp(1e-11, 1).numpy()
    assert p.shape == (90,) and np.allclose(p, ref, rtol=5e-3, atol=1e-6) and abs(p.sum() - 1) < 1e-4
    # critic
    v = uc.infer_critic_all()
    assert np.allclose(v, values, rtol=2e-3, atol=2e-3)
    assert abs(uc.infer_critic_single(5) - values[5]) < 2e-3 * max(1, abs(values[5]))
    # ASSERT_RIGHT_TYPE
    with pytest.raises(engine.EngineError, match="created to infer the critic"):
        uc.infer_policy_all(True)
    with pytest.raises(engine.EngineError, match="created to infer the policy"):
        up.infer_critic_all()
    # a model of another architecture is refused (InferUnit.cpp:31-39)
    with pytest.raises(RuntimeError):
        infer_unit.InferUnit(e, str(tmp_path / "PPO_POLICY.lt"), True, e.obs_size, [64, 64])
    up.close(); uc.close()
