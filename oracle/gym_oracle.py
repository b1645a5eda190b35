"""gym_oracle.py — TEST INFRASTRUCTURE ONLY (oracle).

A plain numpy-float32 restatement of the reference's gym layer for the collection path:
DiscreteAction table, GameState::UpdateFromArena bookkeeping, DefaultOBS / DefaultOBSPadded,
CombinedReward + EventReward + the common rewards, ZeroSumReward, NoTouch/GoalScore terminals.
Every function cites the reference lines it follows (paths under
/root/reference/RLGymPPO_CPP/RLGymSim_CPP/src/RLGymSim_CPP/).

Pinned: tests/test_oracle.py checks it bit-for-bit against tests/golden/gym_*.npz, which were
produced by the unmodified reference (tests/golden/make_golden.py).  The PHYSICS oracle is the
reference itself (oracle/_ref, built by oracle/Makefile) plus tests/golden/tick_*.npz.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32

# CommonValues.h:40-75
BOOST_LOCATIONS = np.array([
    (0, -4240), (-1792, -4184), (1792, -4184), (-3072, -4096), (3072, -4096), (-940, -3308), (940, -3308), (0, -2816),
    (-3584, -2484), (3584, -2484), (-1788, -2300), (1788, -2300), (-2048, -1036), (0, -1024), (2048, -1036), (-3584, 0),
    (-1024, 0), (1024, 0), (3584, 0), (-2048, 1036), (0, 1024), (2048, 1036), (-1788, 2300), (1788, 2300), (-3584, 2484),
    (3584, 2484), (0, 2816), (-940, 3310), (940, 3308), (-3072, 4096), (3072, 4096), (-1792, 4184), (1792, 4184), (0, 4240)],
    dtype=np.float64)
# RocketSim pad order: RLConst.h:215-253 (6 big then 28 small)
PADS_BIG = [(-3584, 0), (3584, 0), (-3072, 4096), (3072, 4096), (-3072, -4096), (3072, -4096)]
PADS_SMALL = [(0, -4240), (-1792, -4184), (1792, -4184), (-940, -3308), (940, -3308), (0, -2816), (-3584, -2484), (3584, -2484),
              (-1788, -2300), (1788, -2300), (-2048, -1036), (0, -1024), (2048, -1036), (-1024, 0), (1024, 0), (-2048, 1036),
              (0, 1024), (2048, 1036), (-1788, 2300), (1788, 2300), (-3584, 2484), (3584, 2484), (0, 2816), (-940, 3308),
              (940, 3308), (-1792, 4184), (1792, 4184), (0, 4240)]


def pad_index_map():
    """GameState.cpp:10-50 _BuildBoostPadIndexMap."""
    rs = np.array(PADS_BIG + PADS_SMALL, dtype=np.float64)
    out = []
    for x, y in BOOST_LOCATIONS:
        d = (rs[:, 0] - x) ** 2 + (rs[:, 1] - y) ** 2
        j = int(np.argmax(d < 10))
        assert d[j] < 10
        out.append(j)
    return np.array(out)


def action_table():
    """Utils/ActionParsers/DiscreteAction.cpp:3-67."""
    acts = []
    for throttle in (-1, 0, 1):
        for steer in (-1, 0, 1):
            for boost in (0, 1):
                for handbrake in (0, 1):
                    if boost == 1 and throttle != 1:
                        continue
                    acts.append((throttle, steer, 0, steer, 0, 0, boost, handbrake))
    for pitch in (-1, 0, 1):
        for yaw in (-1, 0, 1):
            for roll in (-1, 0, 1):
                for jump in (0, 1):
                    for boost in (0, 1):
                        if jump == 1 and yaw != 0:
                            continue
                        if pitch == roll and roll == jump and jump == 0:
                            continue
                        handbrake = int(jump == 1 and (pitch != 0 or yaw != 0 or roll != 0))
                        acts.append((boost, yaw, pitch, yaw, roll, jump, boost, handbrake))
    return np.array(acts, dtype=np.float32)


def _rt(v):
    """Car::SetState -> Car::GetState round trip (RocketSim Car.cpp:10-36): uu * (1/50) stored, * 50 read back."""
    return (np.asarray(v, dtype=np.float32) * f32(f32(1) / f32(50))).astype(np.float32) * f32(50)


def _vlen(v):
    """Vec::Length (RocketSim MathTypes.h:31-41): includes the zero w lane."""
    l2 = f32(f32(f32(v[0] * v[0]) + f32(v[1] * v[1])) + f32(v[2] * v[2])) + f32(0)
    return np.sqrt(l2, dtype=np.float32) if l2 > 0 else f32(0)


def _vnorm(v):
    l = _vlen(v)
    if l > f32(np.finfo(np.float32).eps) * f32(np.finfo(np.float32).eps):
        return np.array([v[0] / l, v[1] / l, v[2] / l], dtype=np.float32)
    return np.zeros(3, dtype=np.float32)



_LIBM = None


def _powf(x, e):
    """glibc powf — the function the reference's SaveBoostReward / TouchBallReward call (same libm as oracle/_ref here)."""
    global _LIBM
    if _LIBM is None:
        import ctypes
        import ctypes.util

        _LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        _LIBM.powf.restype = ctypes.c_float
        _LIBM.powf.argtypes = [ctypes.c_float, ctypes.c_float]
    return f32(_LIBM.powf(float(x), float(e)))


def _vdot(a, b):
    return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2])) + f32(0)


class GymOracle:
    """Holds exactly the cross-step state the reference's Match/Gym keep for the built-in plugins."""

    def __init__(self, cfg, player_order):
        self.cfg = cfg
        self.order = [int(i) - 1 for i in player_order]  # players[i] -> car index
        self.P = len(self.order)
        self.table = action_table()
        self.pad_map = pad_index_map()
        self.terms = [(cfg.reward_terms[i].kind, f32(cfg.reward_terms[i].weight), [f32(x) for x in cfg.reward_terms[i].params])
                      for i in range(cfg.num_reward_terms)]
        self.reset_bookkeeping(0)

    def team(self, ci):
        return (ci & 1) if self.cfg.spawn_opponents else 0

    def reset_bookkeeping(self, tick):
        self.score = [0, 0]
        self.last_tick = 0
        self.steps_since_touch = 0
        self.prev_actions = np.zeros((self.P, 8), dtype=np.float32)
        self.touched = [False] * self.P
        self.memo = None
        self.snap_demoed = [False] * self.P

    # GameState::UpdateFromArena (GameState.cpp:52-104) + PlayerData::UpdateFromCar (PlayerData.cpp:4-33)
    def snapshot(self, cars, ball, tick):
        tick_skip = max(tick - self.last_tick, 0)
        for p, ci in enumerate(self.order):
            c = cars[ci]
            self.touched[p] = bool(c["hit_valid"]) and int(c["hit_tick"]) >= tick - tick_skip
            self.snap_demoed[p] = bool(c["is_demoed"])
        by = _rt(ball["pos"])[1]
        if abs(by) > f32(5124.25) + f32(91.25):  # Math.cpp:3-5
            self.score[1 - (0 if by < 0 else 1)] += 1
        self.last_tick = tick

    def _player_block(self, c, inv):
        """DefaultOBS::AddPlayerToOBS (DefaultOBS.cpp:3-18)."""
        pos = _rt(c["pos"]); vel = _rt(c["vel"]); ang = c["ang_vel"].astype(np.float32)
        fwd = c["rot_forward"].astype(np.float32); up = c["rot_up"].astype(np.float32)
        if inv:
            m = np.array([-1, -1, 1], dtype=np.float32)
            pos, vel, ang, fwd, up = pos * m, vel * m, ang * m, fwd * m, up * m
        pc = np.array([f32(1) / f32(4096), f32(1) / f32(5120), f32(1) / f32(2044)], dtype=np.float32)
        has_flip = (not c["has_double_jumped"]) and (not c["has_flipped"]) and f32(c["air_time_since_jump"]) < f32(1.25)
        return np.concatenate([pos * pc, fwd, up, vel * (f32(1) / f32(2300)), ang * (f32(1) / f32(5.5)),
                               np.array([f32(c["boost"]) / f32(100), f32(bool(c["is_on_ground"])), f32(has_flip), f32(bool(c["is_demoed"]))],
                                        dtype=np.float32)]).astype(np.float32)

    def build_obs(self, cars, ball, pads):
        """DefaultOBS::BuildOBS (DefaultOBS.cpp:20-55) / DefaultOBSPadded::BuildOBS (DefaultOBSPadded.cpp:3-66).
        For the padded builder the teammate / opponent slot blocks are returned UNSHUFFLED (the reference shuffles them
        with the thread RNG): compare those blocks as multisets."""
        rows = []
        pc = np.array([f32(1) / f32(4096), f32(1) / f32(5120), f32(1) / f32(2044)], dtype=np.float32)
        for p, ci in enumerate(self.order):
            inv = self.team(ci) == 1
            bp = _rt(ball["pos"]); bv = _rt(ball["vel"]); ba = ball["ang_vel"].astype(np.float32)
            if inv:
                m = np.array([-1, -1, 1], dtype=np.float32)
                bp, bv, ba = bp * m, bv * m, ba * m
            padv = np.array([f32(bool(pads[self.pad_map[(33 - i) if inv else i]]["is_active"])) for i in range(34)], dtype=np.float32)
            parts = [bp * pc, bv * (f32(1) / f32(2300)), ba * (f32(1) / f32(5.5)), self.prev_actions[p], padv, self._player_block(cars[ci], inv)]
            mates = [self._player_block(cars[cj], inv) for cj in self.order if cj != ci and self.team(cj) == self.team(ci)]
            opps = [self._player_block(cars[cj], inv) for cj in self.order if cj != ci and self.team(cj) != self.team(ci)]
            if self.cfg.obs_kind == 1:
                mp = self.cfg.obs_max_players
                mates += [np.zeros(19, dtype=np.float32)] * (mp - 1 - len(mates))
                opps += [np.zeros(19, dtype=np.float32)] * (mp - len(opps))
            rows.append(np.concatenate(parts + mates + opps).astype(np.float32))
        return np.stack(rows)

    def _event_values(self, cars, p):
        """EventReward::ExtractValues (CommonRewards.cpp:9-24); match counters stay 0 on the eval path."""
        ci = self.order[p]
        t = self.team(ci)
        c = cars[ci]
        return np.array([0, self.score[t], self.score[1 - t], 0, float(self.touched[p]), 0, 0, 0, 0, float(bool(c["is_demoed"])),
                         f32(c["boost"]) / f32(100)], dtype=np.float32)

    def episode_reset(self, cars, ball, tick):
        """Gym::Reset bookkeeping (Gym.cpp:58-66, Match.cpp:4-10, EventReward::Reset)."""
        self.reset_bookkeeping(tick)
        self.snapshot(cars, ball, tick)
        self.steps_since_touch = 0
        self.memo = [self._event_values(cars, p) for p in range(self.P)]

    def _term(self, kind, params, cars, ball, p):
        ci = self.order[p]
        c = cars[ci]
        bpos = _rt(ball["pos"]); bvel = _rt(ball["vel"])
        cpos = _rt(c["pos"]); cvel = _rt(c["vel"])
        if kind == 0:  # EventReward::GetReward (CommonRewards.cpp:32-43)
            nv = self._event_values(cars, p)
            r = f32(0)
            for i in range(11):
                r = f32(r + f32(max(f32(nv[i] - self.memo[p][i]), f32(0)) * params[i]))
            self.memo[p] = nv
            return r
        if kind == 1:  # VelocityPlayerToBallReward (CommonRewards.h:91-98)
            return _vdot(_vnorm(bpos - cpos), cvel / f32(2300))
        if kind == 2:  # VelocityBallToGoalReward (CommonRewards.h:73-88)
            orange = self.team(ci) == 0
            if params[0] != 0:
                orange = not orange
            gz = f32(642.775) / f32(2)
            target = np.array([0, 6000 if orange else -6000, gz], dtype=np.float32)
            return _vdot(_vnorm(target - bpos), bvel / f32(6000))
        if kind == 3:  # FaceBallReward (CommonRewards.h:101-108)
            return _vdot(c["rot_forward"].astype(np.float32), _vnorm(bpos - cpos))
        if kind == 4:  # VelocityReward (CommonRewards.h:52-58)
            return f32(f32(_vlen(cvel) / f32(2300)) * f32(1 - 2 * int(params[0] != 0)))
        if kind == 5:  # SaveBoostReward (CommonRewards.h:61-70); boostFraction = boost / 100 (PlayerData.cpp:32)
            return f32(min(max(_powf(f32(f32(c["boost"]) / f32(100)), f32(params[0])), f32(0)), f32(1)))
        if kind == 6:  # TouchBallReward (CommonRewards.h:110-124)
            if not self.touched[p]:
                return f32(0)
            return _powf(f32(f32(bpos[2] + f32(91.25)) / f32(f32(91.25) * f32(2))), f32(params[0]))
        raise ValueError(kind)

    def rewards(self, cars, ball):
        """CombinedReward::GetAllRewards (CombinedReward.h:36-46) + ZeroSumReward::GetAllRewards (ZeroSumReward.cpp:3-29)."""
        r = [f32(0)] * self.P
        for kind, w, params in self.terms:
            for p in range(self.P):
                r[p] = f32(r[p] + f32(self._term(kind, params, cars, ball, p) * w))
        if self.cfg.zero_sum:
            cnt = [0, 0]; avg = [f32(0), f32(0)]
            for p, ci in enumerate(self.order):
                t = self.team(ci); cnt[t] += 1; avg[t] = f32(avg[t] + r[p])
            for t in range(2):
                avg[t] = f32(avg[t] / f32(max(cnt[t], 1)))
            ts = f32(self.cfg.team_spirit); osc = f32(self.cfg.opponent_scale)
            out = []
            for p, ci in enumerate(self.order):
                t = self.team(ci)
                out.append(f32(f32(f32(r[p] * f32(f32(1) - ts)) + f32(avg[t] * ts)) - f32(avg[1 - t] * osc)))
            r = out
        return np.array(r, dtype=np.float32)

    def done(self, ball):
        """Match::IsDone with [NoTouchCondition, GoalScoreCondition] (NoTouchCondition.h:18-28, GoalScoreCondition.h:9-11)."""
        d = False
        if self.cfg.no_touch_max_steps > 0:
            if any(self.touched):
                self.steps_since_touch = 0
            else:
                self.steps_since_touch += 1
                d = self.steps_since_touch >= self.cfg.no_touch_max_steps
        if not d and self.cfg.goal_score_terminal:
            d = abs(_rt(ball["pos"])[1]) > f32(5124.25) + f32(91.25)
        return d

    def eval(self, cars, ball, pads, tick, actions):
        """Gym::Step (Gym.cpp:68-102) minus physics/event tracker, on an injected state."""
        for p in range(self.P):
            self.prev_actions[p] = 0 if self.snap_demoed[p] else self.table[int(actions[p])]  # Match.cpp:44-52
        self.snapshot(cars, ball, tick)
        obs = self.build_obs(cars, ball, pads)
        d = self.done(ball)
        r = self.rewards(cars, ball)
        return obs, r, d
