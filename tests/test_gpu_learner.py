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
    assert lr.exp.cur_size == 6144
    # the collector now infers with the UPDATED weights
    obs = np.random.default_rng(0).uniform(-1, 1, size=(256, lr.engine.obs_size)).astype(np.float32)
    t_obs = torch.from_numpy(obs).cuda()
    val = torch.empty(256, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    lr.collector.infer(t_obs.data_ptr(), 256, 0, value_ptr=val.data_ptr())
    lr.engine.sync()
    ref = po.mlp_forward(L.mlp_layers_numpy(lr.ppo.value_net), obs)[:, 0]
    assert np.all(np.abs(val.cpu().numpy() - ref) <= 2e-2 + 1e-2 * np.abs(ref))
