"""CPU suite: the oracle is pinned.
 * numpy gym-layer restatement (oracle/gym_oracle.py) == golden fixtures produced by the unmodified reference, bit for bit
 * when the compiled reference (oracle/_ref) is present it still reproduces the committed fixtures (fixtures are not stale)
"""
import os

import numpy as np
import pytest

import common
from oracle import gym_oracle, refsim
from rlgymppo_cpp_b200 import abi


def test_action_table_matches_reference_fixture(golden_dir):
    assert np.array_equal(gym_oracle.action_table(), np.load(os.path.join(golden_dir, "action_table.npy")))


@pytest.mark.parametrize("name,cfg", list(common.gym_cfgs()))
def test_numpy_gym_oracle_bit_exact_vs_reference_golden(name, cfg, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    o = gym_oracle.GymOracle(cfg, g["player_order"])
    n = len(g["tick"])
    stride = 1 if n < 250 else 2  # keep the CPU suite quick; episodes are replayed in full for state carry-over
    for i in range(n):
        cars, ball, pads, tick = g["cars"][i], g["ball"][i], g["pads"][i], int(g["tick"][i])
        if g["first"][i]:
            o.episode_reset(cars, ball, tick)
            obs, r, d = o.build_obs(cars, ball, pads), None, None
        else:
            obs, r, d = o.eval(cars, ball, pads, tick, g["actions"][i])
        if i % stride:
            continue
        assert common.obs_equal(cfg, g["obs"][i], obs), (name, i)
        if r is not None:
            assert common.rewards_equal(name, g["reward"][i], r), (name, i, r, g["reward"][i])
            assert d == bool(g["done"][i]), (name, i)


@pytest.mark.skipif(not refsim.available(), reason="oracle/_ref not built")
def test_compiled_reference_reproduces_tick_fixtures(golden_dir):
    """The recording protocol of make_golden.record (state re-injected before every tick) replayed on a fresh compiled reference:
    bit-identical trajectories, i.e. the fixtures ARE what the reference answers to the recorded states."""
    groups = common.load_tick_file("tick_scenarios_1v1")
    arena = refsim.RefArena(1, True)
    arena.step(None, 1)
    for name in ("free_flight", "jump_flip", "car_hits_ball", "ball_corner_mesh", "car_into_goal"):
        g = groups[name]
        for t in range(len(g["controls"])):
            arena.set_state(g["cars"][t].copy(), g["ball"][t:t + 1].copy(), g["pads"][t].copy(), int(g["tick"][t]))
            arena.step(g["controls"][t].copy(), 1)
            cars, ball, pads, tick = arena.get_state()
            assert np.array_equal(cars["pos"], g["cars"][t + 1]["pos"]) and np.array_equal(cars["vel"], g["cars"][t + 1]["vel"]), (name, t)
            assert np.array_equal(ball["vel"][0], g["ball"][t + 1]["vel"]) and tick == int(g["tick"][t + 1]), (name, t)


@pytest.mark.skipif(not refsim.available(), reason="oracle/_ref not built")
def test_compiled_reference_action_table(golden_dir):
    assert np.array_equal(refsim.action_table(), np.load(os.path.join(golden_dir, "action_table.npy")))
