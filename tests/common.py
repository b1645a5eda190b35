"""Shared by the CPU (hostsim) and GPU (C-ABI engine) parity tests: golden loading + tolerances."""
from __future__ import annotations

import os

import numpy as np

from rlgymppo_cpp_b200 import abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Single-tick tolerances from IDENTICAL start states (uu, uu/s, rad/s). The reference's own notion of
# "close enough" is BallState::Matches(0.8 uu, 0.4 uu/s, 0.02 rad/s) (RocketSim Ball.cpp:12-17); ticks
# without hitbox-vs-world contact are required to be ~1000x tighter than that.
TOL_TIGHT = dict(pos=2e-3, vel=2e-2, ang=2e-4, rot=2e-5)
# A handful of ticks with a deep hitbox contact leave TOL_TIGHT: the penetration-depth search (GJK + EPA on the rounded box,
# rl_epa.h) takes threshold decisions (1e-4 accuracy exit, duplicate-vertex and plane epsilons, closest face) on nearly
# degenerate values, and a different rounding flips single iterations.  Measured over the 13 712 recorded reference ticks
# (profiles/parity_r02_{host,gpu}.json): the host build (IEEE arithmetic) has 1 such tick (0.026 uu/s); the GPU build, whose
# physics is compiled with FMA contraction (build.py), has 6 (worst 0.09 uu, 2.8 uu/s, 0.02 rad/s) — a no-FMA GPU build reproduces
# the host numbers (profiles/r02d_fp_model_ab.txt).  The gate: position inside the reference's own notion of "the same state"
# (BallState::Matches, Ball.cpp:12-17: 0.8 uu, 0.4 uu/s, 0.02 rad/s) by 4x, velocity 10x / angular velocity 2x outside it, on at
# most 0.2 % of a recording's ticks.
TOL_CONTACT = dict(pos=0.2, vel=4.0, ang=0.04, rot=2e-3)
ALLOW_CONTACT_FRAC = 0.002

# every tick test appends its summary here; the -m gpu session writes it to profiles/parity_r02_gpu.json (conftest.py)
PARITY_LOG = {}
_current_fixture = [None]


def record_parity(groups, res):
    name = _current_fixture[0] or "+".join(sorted(groups))[:60]
    PARITY_LOG[name] = dict(res)


def load_tick_file(name):
    _current_fixture[0] = name
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    groups = {}
    for k in z.files:
        g, f = k.split("/")
        groups.setdefault(g, {})[f] = z[k]
    for g in groups.values():
        g["cars"] = g["cars"].view(abi.CAR_DTYPE) if g["cars"].dtype != abi.CAR_DTYPE else g["cars"]
    return groups


def phys_err(ref_cars, ref_ball, got_cars, got_ball):
    e = {}
    d = lambda a, b: float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0
    e["pos"] = max(d(ref_cars["pos"], got_cars["pos"]), d(ref_ball["pos"], got_ball["pos"]))
    e["vel"] = max(d(ref_cars["vel"], got_cars["vel"]), d(ref_ball["vel"], got_ball["vel"]))
    e["ang"] = max(d(ref_cars["ang_vel"], got_cars["ang_vel"]), d(ref_ball["ang_vel"], got_ball["ang_vel"]))
    e["rot"] = max(d(ref_cars["rot_forward"], got_cars["rot_forward"]), d(ref_cars["rot_up"], got_cars["rot_up"]))
    return e


FLAGS = ["is_on_ground", "has_jumped", "has_double_jumped", "has_flipped", "is_flipping", "is_jumping", "is_supersonic",
         "is_auto_flipping", "is_demoed", "hit_valid"]
SCALARS = ["jump_time", "flip_time", "air_time", "air_time_since_jump", "boost", "time_spent_boosting", "supersonic_time",
           "handbrake_val", "auto_flip_timer", "car_contact_cooldown", "demo_respawn_timer"]


def within(e, tol):
    return all(e[k] <= tol[k] for k in tol)


class _TickTally:
    """Classifies every (reference state before, one tick, state after) experiment: inside TOL_TIGHT, inside TOL_CONTACT
    ("loose"), or a failure."""

    def __init__(self):
        self.total = 0
        self.loose = 0
        self.worst = dict(pos=0.0, vel=0.0, ang=0.0, rot=0.0)
        self.worst_loose = dict(pos=0.0, vel=0.0, ang=0.0, rot=0.0)
        self.loose_ticks = []
        self.failures = []

    def add(self, gname, t, g, cars1, ball1, pads1, tick1):
        # Car::Respawn picks one of the team's four respawn spots with the global RNG (Car.cpp:43-56, RLConst.h
        # CAR_RESPAWN_LOCATIONS): the engine's per-arena RNG differs by construction, so on the tick a car respawns its spot
        # is checked for membership and the reference's x is taken for the comparison (y, z, yaw are the same for all four)
        resp = (g["cars"][t]["is_demoed"] != 0) & (g["cars"][t + 1]["is_demoed"] == 0)
        if resp.any():
            cars1 = cars1.copy()
            for ci in np.nonzero(resp)[0]:
                assert min(abs(abs(float(cars1["pos"][ci][0])) - v) for v in (2304.0, 2688.0)) < 1e-2, (gname, t, cars1["pos"][ci])
                cars1["pos"][ci][0] = g["cars"][t + 1]["pos"][ci][0]
        e = phys_err(g["cars"][t + 1], g["ball"][t + 1:t + 2], cars1, ball1)
        self.total += 1
        flags_ok = all(np.array_equal(g["cars"][t + 1][f] != 0, cars1[f] != 0) for f in FLAGS)
        scal_ok = all(np.allclose(g["cars"][t + 1][f], cars1[f], atol=1e-4, rtol=1e-5) for f in SCALARS)
        pads_ok = np.array_equal(g["pads"][t + 1]["is_active"] != 0, pads1["is_active"] != 0)
        if within(e, TOL_TIGHT) and flags_ok and scal_ok and pads_ok and tick1 == int(g["tick"][t + 1]):
            for k in self.worst:
                self.worst[k] = max(self.worst[k], e[k])
        elif within(e, TOL_CONTACT) and flags_ok and scal_ok and pads_ok:
            self.loose += 1
            for k in self.worst_loose:
                self.worst_loose[k] = max(self.worst_loose[k], e[k])
            self.loose_ticks.append((gname, t, {k: round(v, 5) for k, v in e.items()}))
        else:
            self.failures.append((gname, t, e, flags_ok, scal_ok, pads_ok))

    def finish(self, groups, allow_contact_frac, detail):
        assert not self.failures, f"{len(self.failures)} ticks outside the contact tolerance, first: {self.failures[:3]}"
        assert self.loose <= allow_contact_frac * self.total, f"{self.loose}/{self.total} ticks needed the loose contact tolerance: {self.loose_ticks[:5]}"
        res = dict(total=self.total, loose=self.loose, worst_tight=self.worst, worst_loose=self.worst_loose)
        if detail:
            res["loose_ticks"] = self.loose_ticks
        record_parity(groups, res)
        return res


def check_single_tick_run(groups, set_state, tick, get_state, allow_contact_frac=ALLOW_CONTACT_FRAC, detail=False):
    """For every recorded reference tick: inject the reference state BEFORE the tick, run one tick with the recorded
    controls, compare with the reference state AFTER the tick. Returns a summary dict; raises on violations."""
    tally = _TickTally()
    # (the fixtures are recorded on worlds that have stepped before, tick counts >= 1: the cold-start quirk of tick 0 — CarW::solverDt —
    # stays out of the comparison)
    for gname, g in groups.items():
        for t in range(len(g["controls"])):
            set_state(g["cars"][t], g["ball"][t:t + 1], g["pads"][t], int(g["tick"][t]))
            tick(g["controls"][t])
            cars1, ball1, pads1, tick1 = get_state()
            tally.add(gname, t, g, cars1, ball1, pads1, tick1)
    return tally.finish(groups, allow_contact_frac, detail)


def check_single_tick_batch(groups, run_batch, allow_contact_frac=ALLOW_CONTACT_FRAC, detail=False):
    """The same experiment with every recorded tick of the file in its own arena of ONE engine: run_batch(cars [N,P], balls [N],
    pads [N,34], ticks [N], controls [N,P]) injects state i into arena i, steps all arenas one tick with their own controls and
    returns the states after."""
    index = [(gname, t) for gname, g in groups.items() for t in range(len(g["controls"]))]
    g0 = next(iter(groups.values()))
    N, P = len(index), g0["cars"].shape[1]
    # (filled row by row: np.stack re-packs structured dtypes and drops the C struct padding)
    cars = np.zeros((N, P), dtype=abi.CAR_DTYPE)
    balls = np.zeros(N, dtype=abi.BALL_DTYPE)
    pads = np.zeros((N, abi.RLG_NUM_PADS), dtype=abi.PAD_DTYPE)
    ctl = np.zeros((N, P), dtype=abi.CONTROLS_DTYPE)
    for i, (n, t) in enumerate(index):
        cars[i], balls[i], pads[i], ctl[i] = groups[n]["cars"][t], groups[n]["ball"][t], groups[n]["pads"][t], groups[n]["controls"][t]
    ticks = np.array([int(groups[n]["tick"][t]) for n, t in index], dtype=np.int64)
    cars1, balls1, pads1, ticks1 = run_batch(cars, balls, pads, ticks, ctl)
    tally = _TickTally()
    for i, (gname, t) in enumerate(index):
        tally.add(gname, t, groups[gname], cars1[i], balls1[i:i + 1], pads1[i], int(ticks1[i]))
    return tally.finish(groups, allow_contact_frac, detail)


CAR_PRESETS = ((1, "dominus"), (2, "plank"), (3, "breakout"), (4, "hybrid"), (5, "merc"))


def apply_test_mutators(cfg):
    """A MutatorConfig that differs from the default in every field the engine honours (MutatorConfig.h:16-72)."""
    m = abi.default_mutators()
    m.gravity[0], m.gravity[1], m.gravity[2] = 20.0, -15.0, -380.0
    m.car_world_friction, m.car_world_restitution = 0.45, 0.15
    m.ball_max_speed, m.ball_drag, m.ball_world_friction, m.ball_world_restitution = 3500.0, 0.06, 0.5, 0.75
    m.jump_accel, m.jump_immediate_force = 1800.0, 350.0
    m.boost_accel_ground, m.boost_accel_air, m.boost_used_per_second = 1400.0, 1500.0, 20.0
    m.respawn_delay, m.bump_cooldown_time = 1.0, 0.1
    m.boost_pad_cooldown_big, m.boost_pad_cooldown_small = 3.0, 1.5
    m.car_spawn_boost_amount = 60.0
    m.ball_hit_extra_force_scale, m.bump_force_scale = 1.6, 0.5
    m.unlimited_flips, m.unlimited_double_jumps = 1, 1
    m.demo_mode, m.enable_team_demos = 1, 1  # ON_CONTACT
    m.goal_base_threshold_y = 5000.0
    cfg.mutators = m
    cfg.mutators_set = 1
    return cfg


def apply_ball_mutators(cfg):
    """The three mass / size fields of MutatorConfig (MutatorConfig.h:21,29,59): a heavier, larger ball (Ball.cpp:74-91 builds the ball
    from them) and a carMass that the reference's Gym never applies (Gym.cpp:40-49 sets the mutators before the cars exist)."""
    m = abi.default_mutators()
    m.ball_mass, m.ball_radius, m.car_mass = 45.0, 100.0, 260.0  # (the reference's broadphase refuses balls above ~102 uu)
    cfg.mutators = m
    cfg.mutators_set = 1
    return cfg


def gym_cfgs():
    c = abi.default_cfg(1, 1)
    for k in range(11):
        c.reward_terms[3].params[k] = 0.1 * (k + 1)
    yield "gym_1v1_default", c
    c = abi.default_cfg(1, 2)
    c.zero_sum = 1; c.team_spirit = 0.3; c.num_reward_terms = 5
    c.reward_terms[4].kind = abi.RLG_REW_VELOCITY; c.reward_terms[4].weight = 0.25
    yield "gym_2v2_zerosum", c
    c = abi.default_cfg(1, 3)
    c.obs_kind = abi.RLG_OBS_PADDED; c.obs_max_players = 3
    yield "gym_3v3_padded", c
    c = abi.default_cfg(1, 2)
    c.obs_kind = abi.RLG_OBS_PADDED; c.obs_max_players = 3
    c.zero_sum = 1; c.team_spirit = 0.3; c.state_setter = abi.RLG_SETTER_KICKOFF
    yield "gym_2v2_padded_zerosum_kickoff", c
    c = extra_rewards_cfg()
    yield "gym_1v1_extra_rewards", c


def extra_rewards_cfg():
    """cfg-1 rewards + SaveBoostReward(0.5), SaveBoostReward(0.37), TouchBallReward(0.8), TouchBallReward(0) (the powf rewards:
    compared within 1 ulp, see REWARD_ULPS)."""
    c = abi.default_cfg(1, 1)
    c.num_reward_terms = 8
    for i, (kind, w, p0) in enumerate([(abi.RLG_REW_SAVE_BOOST, 0.3, 0.5), (abi.RLG_REW_SAVE_BOOST, 0.2, 0.37), (abi.RLG_REW_TOUCH_BALL, 2.0, 0.8),
                                        (abi.RLG_REW_TOUCH_BALL, 1.0, 0.0)]):
        c.reward_terms[4 + i].kind = kind; c.reward_terms[4 + i].weight = w; c.reward_terms[4 + i].params[0] = p0
    c.reward_terms[3].params[4] = 0.5  # EventReward touch weight: touches show up in two places
    return c


# rewards are bit-exact, except configurations with a powf reward (SaveBoostReward / TouchBallReward): every such term may be 1 ulp
# off glibc's powf, and the weighted sum carries it: |delta| <= REWARD_ULPS ulps of the largest term magnitude (~3)
REWARD_ULPS = {"gym_1v1_extra_rewards": 8}


def rewards_equal(name, ref, got):
    ulps = REWARD_ULPS.get(name, 0)
    if ulps == 0:
        return np.array_equal(np.asarray(ref, np.float32).view(np.uint32), np.asarray(got, np.float32).view(np.uint32))
    scale = np.maximum(np.abs(np.asarray(ref, np.float64)), 1.0)
    return bool(np.all(np.abs(np.asarray(ref, np.float64) - np.asarray(got, np.float64)) <= ulps * scale * 2.0 ** -23))


def obs_equal(cfg, ref, got):
    """bit-exact; for the padded builder the shuffled teammate/opponent slot blocks compare as multisets."""
    if cfg.obs_kind == abi.RLG_OBS_DEFAULT:
        return np.array_equal(ref.view(np.uint32), got.view(np.uint32))
    mp = cfg.obs_max_players
    if not np.array_equal(ref[:, :70].view(np.uint32), got[:, :70].view(np.uint32)):
        return False
    for p in range(ref.shape[0]):
        for lo, n in ((70, mp - 1), (70 + 19 * (mp - 1), mp)):
            a = sorted(bytes(x) for x in ref[p, lo:lo + 19 * n].reshape(n, 19))
            b = sorted(bytes(x) for x in got[p, lo:lo + 19 * n].reshape(n, 19))
            if a != b:
                return False
    return True


# ---- state setters: distributions against the live reference --------------------------------------------------------------
def setter_features(cars, ball):
    """Scalar features of one reset state (cars [P] CAR_DTYPE, ball [1] BALL_DTYPE) whose distributions RandomState / KickoffState
    define (RandomState.cpp:8-62, Arena.cpp:112-216)."""
    f = {}
    b = ball[0] if ball.shape else ball
    for i, ax in enumerate("xyz"):
        f["ball_pos_" + ax] = [float(b["pos"][i])]
        f["ball_angvel_" + ax] = [float(b["ang_vel"][i])]
    v = np.asarray(b["vel"], np.float64)
    sp = float(np.linalg.norm(v))
    f["ball_speed"] = [sp]
    f["ball_vel_dir_z"] = [float(v[2] / sp)] if sp > 1e-6 else []
    f["car_pos_x"] = [float(c["pos"][0]) for c in cars]
    f["car_pos_y"] = [float(c["pos"][1]) for c in cars]
    f["car_yaw"] = [float(np.arctan2(c["rot_forward"][1], c["rot_forward"][0])) for c in cars]
    f["car_speed"] = [float(np.linalg.norm(np.asarray(c["vel"], np.float64))) for c in cars]
    f["car_vel_heading"] = [float(np.arctan2(c["vel"][1], c["vel"][0])) for c in cars if np.linalg.norm(c["vel"]) > 1e-6]
    f["car_boost"] = [float(c["boost"]) for c in cars]
    return f


def ks_two_sample(a, b):
    """Two-sample Kolmogorov-Smirnov statistic D and its asymptotic critical value at alpha = 1e-4."""
    a, b = np.sort(np.asarray(a, np.float64)), np.sort(np.asarray(b, np.float64))
    allv = np.concatenate([a, b])
    d = float(np.max(np.abs(np.searchsorted(a, allv, side="right") / len(a) - np.searchsorted(b, allv, side="right") / len(b))))
    crit = float(np.sqrt(-0.5 * np.log(1e-4 / 2)) * np.sqrt((len(a) + len(b)) / (len(a) * len(b))))
    return d, crit


def compare_setter_samples(ours, ref):
    """ours / ref: lists of (cars, ball, pads) reset states.  Every feature's distribution must pass a KS test against the
    reference's at alpha = 1e-4 per feature (14 features: family-wise < 0.2 %), the deterministic parts must be equal."""
    fo, fr = {}, {}
    for acc, samples in ((fo, ours), (fr, ref)):
        for cars, ball, pads in samples:
            for k, v in setter_features(cars, ball).items():
                acc.setdefault(k, []).extend(v)
            assert np.all(pads["is_active"] != 0)
            assert np.allclose(cars["pos"][:, 2], 17.0) and np.all(cars["is_on_ground"] != 0)  # carsOnGround = true (examplemain.cpp:85)
            assert np.all(cars["vel"][:, 2] == 0) and np.all(cars["ang_vel"] == 0)
            assert np.allclose(cars["rot_up"], [0, 0, 1], atol=1e-6)
    report = {}
    for k in fr:
        d, crit = ks_two_sample(fo[k], fr[k])
        report[k] = (round(d, 4), round(crit, 4), len(fo[k]), len(fr[k]))
        assert d < crit, (k, report[k])
    return report
