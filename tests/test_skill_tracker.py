"""SkillTracker host logic (rlgymppo_cpp_b200/skill_tracker.py) against the reference's rules
(RLGymPPO_CPP/src/private/RLGymPPO_CPP/Util/SkillTracker.cpp:72-86 UpdateRatings, :88-157 RunThread, :159-257 RunGames),
on CPU with a scripted environment; the device wiring is covered by tests/test_gpu_collector.py::test_skill_tracker_on_engine."""
import numpy as np

from rlgymppo_cpp_b200 import skill_tracker as stm


def test_update_ratings_known_answers():
    # equal ratings: expected = 1 / (10^0 + 1) = 0.5 -> +-ratingInc/2
    w, l = {"1v1": 1000.0}, {"1v1": 1000.0}
    stm.update_ratings(w, l, 5.0, "1v1")
    assert w["1v1"] == 1002.5 and l["1v1"] == 997.5
    # winner 400 points ahead: expDelta = -1, expected = 1 / (0.1 + 1) = 10/11 -> winner gains 5/11
    w, l = {"": 1400.0}, {"": 1000.0}
    stm.update_ratings(w, l, 5.0, "")
    f = np.float32
    exp = f(1) / (f(np.power(f(10), f(-1), dtype=f)) + f(1))
    assert w[""] == float(f(1400) + f(5) * (f(1) - exp)) and l[""] == float(f(1000) + f(5) * (exp - f(1)))
    assert abs(w[""] - (1400 + 5 / 11)) < 1e-3
    # zero-sum
    assert abs((w[""] - 1400) + (l[""] - 1000)) < 1e-4


def test_mode_names():
    assert stm.mode_name(1, True) == "1v1" and stm.mode_name(3, True) == "3v3" and stm.mode_name(2, False) == "2v0"


def test_select_actions_team_swap():
    teams = np.array([0, 1, 0, 1])  # 2v2 player slots: blue, orange, blue, orange
    cur = np.arange(8, dtype=np.int32)          # arena 0 rows 0..3, arena 1 rows 4..7
    old = {0: cur + 100, 1: cur + 200}
    out = stm.select_actions(teams, np.array([False, True]), np.array([0, 1]), cur, old, 4)
    # arena 0 not swapped: blue = current, orange = old[0]; arena 1 swapped: blue = old[1], orange = current
    assert out.tolist() == [0, 101, 2, 103, 204, 5, 206, 7]


def _scripted(cfg, goals_by_step):
    """A tracker over a fake env: goals_by_step[t] = list of (arena, +1 | -1)."""
    st = stm.SkillTracker(cfg, team_size=1, spawn_opponents=True, tick_skip=8, seed=3)
    st.teams = np.array([0, 1])
    t = [0]
    log = {"infer": [], "resets": 0}

    def infer(weights):
        log["infer"].append(weights)
        return np.zeros(cfg.numEnvs * 2, dtype=np.int32)

    def step(actions):
        scored = np.zeros(cfg.numEnvs, dtype=np.int32)
        for a, s in goals_by_step.get(t[0], []):
            scored[a] = s
        t[0] += 1
        return (scored != 0).astype(np.uint8), scored

    def reset_all():
        log["resets"] += 1

    st.infer_fn, st.step_fn, st.reset_all_fn = infer, step, reset_all
    return st, log


def test_run_games_bookkeeping_and_elo():
    cfg = stm.SkillTrackerConfig(enabled=True, numEnvs=2, simTime=2 * 8 * 3 / 120, updateInterval=2, timestepsPerVersion=1000, maxVersions=2)
    # simTime / numEnvs * 120 / tickSkip = 3 steps per RunGames
    st, log = _scripted(cfg, {0: [(0, +1)], 2: [(1, -1)]})
    st.team_swap[:] = [False, True]
    r = st.run_games("w0", 400)
    # first call: run_counter 0 -> plays; startWithVersion froze "w0" as version 0
    assert r is not None and st.old_policies == ["w0"] and len(st.old_ratings) == 1
    # arena 0: blue scored, not swapped -> current wins; arena 1: orange scored, swapped -> orange = current wins
    assert st.goals == 2
    w, l = {"1v1": 1000.0}, {"1v1": 1000.0}
    stm.update_ratings(w, l, 5.0, "1v1")
    stm.update_ratings(w, l, 5.0, "1v1")
    assert st.cur_rating == w and st.old_ratings[0] == l
    assert len(log["infer"]) == 3 * 2  # current + one old version per step
    # second call is skipped by updateInterval: RunGames returns before it counts the timesteps (:162-167)
    assert st.run_games("w1", 400) is None and st.timesteps_since_version == 400
    # third call plays again and crosses timestepsPerVersion (400 + 700): all games re-set, a new version is frozen
    assert st.run_games("w2", 700) is not None
    assert log["resets"] == 1 and st.old_policies == ["w0", "w2"] and st.timesteps_since_version == 0
    st.run_counter = 0
    st.run_games("w3", 5000)
    assert st.old_policies == ["w2", "w3"] and len(st.old_ratings) == 2 and st.old_index.max() <= 1  # maxVersions


def test_episode_end_redraws_side_and_opponent():
    cfg = stm.SkillTrackerConfig(enabled=True, numEnvs=1, simTime=8 * 50 / 120, updateInterval=1, timestepsPerVersion=10 ** 9)
    st, _ = _scripted(cfg, {i: [(0, +1)] for i in range(50)})
    st.append_old_policy("a", st.cur_rating); st.append_old_policy("b", st.cur_rating); st.append_old_policy("c", st.cur_rating)
    seen = set()
    orig = st._reset_game

    def spy(a, n):
        orig(a, n)
        seen.add((bool(st.team_swap[a]), int(st.old_index[a])))

    st._reset_game = spy
    st.run_games("cur", 0)
    assert {s for s, _ in seen} == {True, False} and {i for _, i in seen} == {0, 1, 2}
