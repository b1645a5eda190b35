"""SkillTracker on the device engine: ELO of the current policy against frozen older versions.

Reference: /root/reference/RLGymPPO_CPP/src/private/RLGymPPO_CPP/Util/SkillTracker.cpp:27-257 (+ SkillTracker.h:12-63,
public/RLGymPPO_CPP/Util/SkillTrackerConfig.h:7-44).  What the reference does with ``numEnvs`` GameInsts on host threads is
done here with ONE small engine holding ``numEnvs`` arenas:

* every eval arena plays current policy vs ONE old version (``Game.oldPolicyIndex``), sides swapped at random per episode
  (``Game.teamSwap``); both are re-drawn when the arena's episode ends (SkillTracker.cpp:151-152, SkillTracker.h:24-27);
* rewards are a dummy (zero reward terms: SkillTracker.cpp:11-17,51), states are kickoffs when ``kickoffStatesOnly``;
* per step the obs matrix of ALL eval arenas goes through the tcgen05 inference kernel once per policy version in play
  (deterministic = argmax, ``GetAction(obs, 1)`` :118-119) and a per-row select picks the action of the policy that owns
  the row's team;
* a step whose state has the ball behind a goal line (``Math::IsBallScored``, G/Math.cpp:3-5) is a goal for the policy on
  the scoring side (ball.y > 0 => blue scored, :133-134) and moves the two ratings by the ELO rule of ``UpdateRatings``
  (:72-86) in float32; the reference serialises these updates with a mutex in thread-arrival order, here they are applied
  in arena order;
* ``RunGames`` bookkeeping (updateInterval, startWithVersion, timestepsPerVersion, maxVersions: :159-257) is kept as is.

The eval pool lives on rank 0 only.  Ratings are per mode name ("1v1", "2v2", ... or "" when
``perModeRatings`` is off); one engine has one mode.
"""
from __future__ import annotations

import copy
import dataclasses
from typing import Dict, List, Optional

import numpy as np


@dataclasses.dataclass
class SkillTrackerConfig:  # SkillTrackerConfig.h:7-44 (envCreateFunc / stepCallback are host-plugin hooks)
    enabled: bool = False
    numEnvs: int = 4
    simTime: float = 60.0
    updateInterval: int = 4
    timestepsPerVersion: int = 50 * 1000 * 1000
    maxVersions: int = 4
    numThreads: int = 8  # kept for source compatibility: the eval arenas step in one launch
    perModeRatings: bool = True
    loadOldVersionsFromCheckpoints: bool = True
    startWithVersion: bool = True
    kickoffStatesOnly: bool = True
    ratingInc: float = 5.0
    initialRating: float = 1000.0


def update_ratings(winner: Dict[str, float], loser: Dict[str, float], rating_inc: float, mode: str, update_winner=True, update_loser=True):
    """SkillTracker::UpdateRatings (SkillTracker.cpp:72-86), float32 like the reference (powf)."""
    f = np.float32
    if mode not in winner or mode not in loser:
        raise RuntimeError("SkillTracker::UpdateRatings(): missing mode " + repr(mode))  # RG_ASSERT
    exp_delta = f(f(f(loser[mode]) - f(winner[mode])) / f(400))
    expected = f(f(1) / f(np.power(f(10), exp_delta, dtype=f) + f(1)))
    if update_winner:
        winner[mode] = float(f(f(winner[mode]) + f(f(rating_inc) * f(f(1) - expected))))
    if update_loser:
        loser[mode] = float(f(f(loser[mode]) + f(f(rating_inc) * f(expected - f(1)))))


def mode_name(team_size: int, spawn_opponents: bool) -> str:
    """ModeNameFromGameInst (SkillTracker.cpp:19-25)."""
    return f"{team_size}v{team_size}" if spawn_opponents else f"{team_size}v0"


def select_actions(teams: np.ndarray, team_swap: np.ndarray, old_index: np.ndarray, cur_actions: np.ndarray,
                   old_actions: Dict[int, np.ndarray], P: int) -> np.ndarray:
    """Per-row merge of the two policies' actions (SkillTracker.cpp:112-130): blue rows take the current policy's action
    unless the arena is team-swapped, orange rows the other way round.  teams: [P] 0 = blue / 1 = orange in player order;
    team_swap / old_index: [A]; cur_actions / old_actions[i]: [A * P]."""
    A = len(team_swap)
    out = np.empty(A * P, dtype=np.int32)
    for a in range(A):
        old = old_actions[int(old_index[a])]
        for p in range(P):
            blue = teams[p] == 0
            use_cur = blue != bool(team_swap[a])
            out[a * P + p] = cur_actions[a * P + p] if use_cur else old[a * P + p]
    return out


class SkillTracker:
    """Host bookkeeping is engine-agnostic (``step_fn`` / ``infer_fn`` injected) so the CPU suite can test it without a GPU;
    ``SkillTracker.on_engine`` wires it to the device engine + collector."""

    def __init__(self, config: SkillTrackerConfig, team_size: int, spawn_opponents: bool, tick_skip: int, seed: int = 0):
        assert config.numEnvs > 0 and config.timestepsPerVersion >= 0 and config.maxVersions > 0 and config.simTime > 0  # :30-34
        self.config = config
        self.tick_skip = tick_skip
        self.mode = mode_name(team_size, spawn_opponents) if config.perModeRatings else ""
        self.cur_rating: Dict[str, float] = {self.mode: float(np.float32(config.initialRating))}
        self.old_policies: List[object] = []   # frozen weight sets, oldest first
        self.old_ratings: List[Dict[str, float]] = []
        self.timesteps_since_version = 0
        self.run_counter = 0
        self.rng = np.random.default_rng(seed)
        self.team_swap = np.zeros(config.numEnvs, dtype=bool)
        self.old_index = np.zeros(config.numEnvs, dtype=np.int64)
        for a in range(config.numEnvs):
            self._reset_game(a, 1)  # Game(gameInst, 1) (:66)
        self.goals = 0
        # wired by on_engine() or by a test
        self.reset_all_fn = None   # () -> None: GameInst::Start for every eval arena
        self.step_fn = None        # (actions [A*P] i32) -> (done [A] u8, scored [A] i32: +1 blue / -1 orange / 0)
        self.infer_fn = None       # (weights) -> actions [A*P] i32 for the CURRENT obs of the eval arenas
        self.teams = None          # [P] team of each player slot

    def _reset_game(self, a: int, num_policies: int):
        """Game::Reset (SkillTracker.h:24-27)."""
        self.team_swap[a] = self.rng.random() > 0.5
        self.old_index[a] = int(self.rng.integers(0, max(num_policies, 1)))

    def append_old_policy(self, weights, rating: Dict[str, float]):
        self.old_policies.append(copy.deepcopy(weights))
        self.old_ratings.append(dict(rating))

    def run_games(self, cur_weights, timesteps_delta: int) -> Optional[Dict[str, float]]:
        """SkillTracker::RunGames (:159-257).  Returns the ratings when games were played, else None."""
        cfg = self.config
        skip = self.run_counter % cfg.updateInterval != 0
        self.run_counter += 1
        if skip:
            return None
        if not self.old_policies and cfg.startWithVersion:
            self.append_old_policy(cur_weights, self.cur_rating)
        played = None
        if self.old_policies:
            # every game gets simTime / numGames seconds (:181, :98-101)
            num_steps = int((cfg.simTime / cfg.numEnvs) * 120 / self.tick_skip)
            if num_steps <= 0:
                raise RuntimeError("RLGPC::SkillTracker RunThread(): simTime is too low for the number of games, there is not enough time per game to step")
            P = len(self.teams)
            for _ in range(num_steps):
                cur_actions = self.infer_fn(cur_weights)
                needed = sorted(set(int(i) for i in self.old_index))
                old_actions = {i: self.infer_fn(self.old_policies[i]) for i in needed}
                actions = select_actions(self.teams, self.team_swap, self.old_index, cur_actions, old_actions, P)
                done, scored = self.step_fn(actions)
                for a in range(cfg.numEnvs):
                    if scored[a] != 0:  # Math::IsBallScored(stepResult.state.ball.pos): +1 blue scored (ball.y > 0), -1 orange
                        blue_scored = scored[a] > 0
                        blue_is_cur = not self.team_swap[a]
                        old = self.old_ratings[int(self.old_index[a])]
                        if blue_scored == blue_is_cur:
                            update_ratings(self.cur_rating, old, cfg.ratingInc, self.mode)
                        else:
                            update_ratings(old, self.cur_rating, cfg.ratingInc, self.mode)
                        self.goals += 1
                    if done[a]:
                        self._reset_game(a, len(self.old_policies))
            played = dict(self.cur_rating)
        self.timesteps_since_version += int(timesteps_delta)
        if self.timesteps_since_version >= cfg.timestepsPerVersion:
            self.reset_all_fn()  # "Reset all games" (:232-234)
            self.timesteps_since_version = 0
            self.append_old_policy(cur_weights, self.cur_rating)
            if len(self.old_policies) > cfg.maxVersions:
                self.old_policies.pop(0)
                self.old_ratings.pop(0)
                # the reference keeps stale Game.oldPolicyIndex values until the next Game::Reset; an index past the end
                # would read freed memory there, so clamp
                np.minimum(self.old_index, len(self.old_policies) - 1, out=self.old_index)
        return played

    # ---- device wiring -------------------------------------------------------------------------------------------
    @classmethod
    def on_engine(cls, config: SkillTrackerConfig, train_engine_cfg, policy_hidden, device_index: int = 0, seed: int = 0, mesh_blobs=None,
                  action_table=None):
        """Eval pool = a second engine with numEnvs arenas of the training configuration (same mode, obs builder,
        terminal conditions), dummy rewards, kickoff states if configured; inference through one deterministic collector."""
        import torch

        from . import abi, collector, engine

        ecfg = type(train_engine_cfg).from_buffer_copy(train_engine_cfg)  # ctypes rlg_engine_cfg
        ecfg.num_arenas = config.numEnvs
        ecfg.device = device_index
        ecfg.arena_id_base = 1 << 24  # RNG streams apart from the training arenas
        ecfg.num_reward_terms = 0     # DummyReward
        ecfg.zero_sum = 0
        if config.kickoffStatesOnly or ecfg.state_setter == abi.RLG_SETTER_HOST:
            # a host StateSetter of the training pool is not wired into the eval pool: its games start from kickoffs
            ecfg.state_setter = abi.RLG_SETTER_KICKOFF
        e = engine.Engine(ecfg, mesh_blobs=mesh_blobs)
        if action_table is not None:  # the training pool's user ActionParser
            e.set_action_table(action_table)
        col = collector.Collector(e, tuple(policy_hidden), tuple(policy_hidden), max_steps=1, seed=seed, deterministic=True)
        st = cls(config, int(ecfg.team_size), bool(ecfg.spawn_opponents), int(ecfg.tick_skip), seed)
        st.engine, st.collector = e, col
        order = [int(i) - 1 for i in e.player_order()]
        st.teams = np.array([(ci & 1) if ecfg.spawn_opponents else 0 for ci in order], dtype=np.int32)  # car_team (rl_state.h)
        A, P = e.A, e.P
        act = torch.empty(A * P, dtype=torch.int32, device=f"cuda:{device_index}")
        counter = [0]

        # one collector (= one packed TF32 weight set on the device) per policy version in play: a version is packed once,
        # when it first plays, not on every step.  The cache holds a reference to the weight object, so its id stays unique.
        packed = {}
        spare = [col]

        def collector_for(weights):
            hit = packed.get(id(weights))
            if hit is not None:
                return hit[1]
            if len(packed) > config.maxVersions + 2:  # versions that left the pool (and previous current policies)
                live = {id(w) for w in st.old_policies}
                for k in [k for k in packed if k not in live]:
                    spare.append(packed.pop(k)[1])
            c = spare.pop() if spare else collector.Collector(e, tuple(policy_hidden), tuple(policy_hidden), max_steps=1, seed=seed, deterministic=True)
            if c.weights[1] is None:
                c.init_default(seed)  # the critic is never run
            c.set_weights(0, weights)
            packed[id(weights)] = (weights, c)
            return c

        def infer(weights):
            c = collector_for(weights)
            obs_ptr, _, _ = e.output_ptrs()
            c.infer(obs_ptr, A * P, counter[0], action_ptr=act.data_ptr())
            counter[0] += 1
            e.sync()
            return act.cpu().numpy().copy()

        prev_lines = np.zeros((A, 2), dtype=np.int32)

        def step(actions):
            # the state the reference inspects is the step's snapshot (taken after its first tick); that snapshot bumps
            # GameState::scoreLine exactly when Math::IsBallScored holds, so: step without the auto-reset, diff the score
            # lines, then re-set the finished arenas (== GameInst::Step's reset)
            t = torch.from_numpy(np.ascontiguousarray(actions, dtype=np.int32)).to(act.device)
            e.step_device(t.data_ptr(), auto_reset=False)
            _, _, done = e.read_outputs()
            lines = e.score_lines()
            delta = lines - prev_lines
            scored = np.where(delta[:, 0] > 0, 1, np.where(delta[:, 1] > 0, -1, 0)).astype(np.int32)
            prev_lines[:] = lines
            if done.any():
                e.reset(mask=done.astype(np.uint8))
                prev_lines[done != 0] = 0
            return done, scored

        def reset_all():
            e.reset()
            prev_lines[:] = 0

        st.infer_fn, st.step_fn, st.reset_all_fn = infer, step, reset_all
        e.reset()
        return st
