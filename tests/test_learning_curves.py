"""The learning-curve A/B evidence (profiles/learning_curves_r02.json, written by tools/learning_curves.py on the B200 box): the
SAME learner trained on the device engine and on the reference's CPU gyms (oracle/_ref), 5 seeds each, examplemain settings.
north_star asks for learning curves that are statistically indistinguishable; the reference's own correctness claim is a training
run (/root/reference/README.md:30).  This test reads the committed evidence and checks the overlap criterion."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_engine_and_reference_learning_curves_overlap():
    j = json.load(open(os.path.join(ROOT, "profiles", "learning_curves_r02.json")))
    eng, ref, cmp_ = j["arms"]["engine"], j["arms"]["reference"], j["comparison"]
    assert eng["seeds"] >= 5 and ref["seeds"] >= 5
    assert eng["timesteps_per_seed"] >= 50e6 and cmp_["timesteps_compared"] >= 20e6
    for k, v in cmp_["metrics"].items():
        z = np.abs(np.array(v["block_z"]))
        # 19 blocks of 10 iterations x 8 metrics: every block mean within 3 standard errors of the difference, and per iteration
        # (5 + 5 seeds: heavy-tailed) at least 95 % within 3 SE
        assert z.max() < 3.0, (k, z.max())
        assert v["frac_iterations_within_3se"] >= 0.95, (k, v["frac_iterations_within_3se"])
    # and the policy does learn on the engine: the mean step reward rises from the random-policy level, the entropy falls
    s = eng["summary"]
    r, h = np.array(s["step_reward"]["mean"]), np.array(s["entropy"]["mean"])
    assert r[-20:].mean() > r[:20].mean() + 0.05 and h[-20:].mean() < h[:20].mean() - 0.1
