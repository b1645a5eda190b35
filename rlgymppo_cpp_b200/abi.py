"""ctypes mirrors of the structs in include/rlgym_b200.h (the C-ABI boundary).

Both the CUDA engine (csrc/librlgym_b200.so) and the test-only reference harness
(oracle/_ref/librlref.so) speak these structs, so parity tests move states between
the two without translation.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

RLG_MAX_CARS = 6
RLG_NUM_PADS = 34
RLG_NUM_ACTIONS = 90
RLG_MAX_REWARD_TERMS = 8

RLG_OBS_DEFAULT, RLG_OBS_PADDED = 0, 1
RLG_SETTER_KICKOFF, RLG_SETTER_RANDOM, RLG_SETTER_HOST = 0, 1, 2
RLG_CAR_OCTANE, RLG_CAR_DOMINUS, RLG_CAR_PLANK, RLG_CAR_BREAKOUT, RLG_CAR_HYBRID, RLG_CAR_MERC = range(6)
RLG_REW_EVENT, RLG_REW_VEL_PLAYER_TO_BALL, RLG_REW_VEL_BALL_TO_GOAL, RLG_REW_FACE_BALL, RLG_REW_VELOCITY, RLG_REW_SAVE_BOOST, RLG_REW_TOUCH_BALL = range(7)

F3 = C.c_float * 3
F4 = C.c_float * 4
I4 = C.c_int32 * 4


class Controls(C.Structure):
    _fields_ = [
        ("throttle", C.c_float), ("steer", C.c_float), ("pitch", C.c_float), ("yaw", C.c_float),
        ("roll", C.c_float), ("jump", C.c_int32), ("boost", C.c_int32), ("handbrake", C.c_int32),
    ]


class CarState(C.Structure):
    _fields_ = [
        ("pos", F3), ("rot_forward", F3), ("rot_right", F3), ("rot_up", F3), ("vel", F3), ("ang_vel", F3),
        ("is_on_ground", C.c_int32), ("wheels_with_contact", I4),
        ("has_jumped", C.c_int32), ("has_double_jumped", C.c_int32), ("has_flipped", C.c_int32),
        ("flip_rel_torque", F3), ("jump_time", C.c_float), ("flip_time", C.c_float),
        ("is_flipping", C.c_int32), ("is_jumping", C.c_int32),
        ("air_time", C.c_float), ("air_time_since_jump", C.c_float),
        ("boost", C.c_float), ("time_spent_boosting", C.c_float),
        ("is_supersonic", C.c_int32), ("supersonic_time", C.c_float), ("handbrake_val", C.c_float),
        ("is_auto_flipping", C.c_int32), ("auto_flip_timer", C.c_float), ("auto_flip_torque_scale", C.c_float),
        ("world_contact_has", C.c_int32), ("world_contact_normal", F3),
        ("car_contact_other_id", C.c_int32), ("car_contact_cooldown", C.c_float),
        ("is_demoed", C.c_int32), ("demo_respawn_timer", C.c_float),
        ("hit_valid", C.c_int32), ("hit_rel_pos_on_ball", F3), ("hit_ball_pos", F3), ("hit_extra_vel", F3),
        ("hit_tick", C.c_int64), ("hit_extra_tick", C.c_int64),
        ("last_controls", Controls),
        ("wheel_steer_angle", C.c_float), ("wheel_engine_force", C.c_float), ("wheel_brake", C.c_float),
        ("wheel_lat_friction", F4), ("wheel_long_friction", F4), ("wheel_extra_pushback", F4),
        ("car_id", C.c_int32), ("team", C.c_int32),
    ]


class MetricsHost(C.Structure):
    """rlg_metrics_host: GameInst's avgStepRew / avgEpRew summed over the arenas (GameInst.cpp:13-31, ThreadAgentManager.cpp:82-92)."""
    _fields_ = [("avg_step_reward", C.c_float), ("avg_episode_reward", C.c_float), ("step_reward_total", C.c_float),
                ("episode_reward_total", C.c_float), ("step_reward_count", C.c_uint64), ("episode_count", C.c_uint64),
                ("total_steps", C.c_uint64)]


class BallState(C.Structure):
    _fields_ = [("pos", F3), ("vel", F3), ("ang_vel", F3)]


class PadState(C.Structure):
    _fields_ = [("is_active", C.c_int32), ("cooldown", C.c_float), ("prev_locked_car_id", C.c_int32)]


class RewardTerm(C.Structure):
    _fields_ = [("kind", C.c_int32), ("weight", C.c_float), ("params", C.c_float * 11)]


class Mutators(C.Structure):
    """rlg_mutators (MutatorConfig.h:16-72)."""
    _fields_ = [("gravity", F3), ("car_mass", C.c_float), ("car_world_friction", C.c_float), ("car_world_restitution", C.c_float),
                ("ball_mass", C.c_float), ("ball_max_speed", C.c_float), ("ball_drag", C.c_float), ("ball_world_friction", C.c_float),
                ("ball_world_restitution", C.c_float), ("jump_accel", C.c_float), ("jump_immediate_force", C.c_float),
                ("boost_accel_ground", C.c_float), ("boost_accel_air", C.c_float), ("boost_used_per_second", C.c_float),
                ("respawn_delay", C.c_float), ("bump_cooldown_time", C.c_float), ("boost_pad_cooldown_big", C.c_float),
                ("boost_pad_cooldown_small", C.c_float), ("car_spawn_boost_amount", C.c_float), ("ball_hit_extra_force_scale", C.c_float),
                ("bump_force_scale", C.c_float), ("ball_radius", C.c_float), ("unlimited_flips", C.c_int32), ("unlimited_double_jumps", C.c_int32),
                ("demo_mode", C.c_int32), ("enable_team_demos", C.c_int32), ("goal_base_threshold_y", C.c_float)]


def default_mutators() -> "Mutators":
    """MutatorConfig(GameMode::SOCCAR) (MutatorConfig.cpp; RLConst.h)."""
    m = Mutators()
    m.gravity[0], m.gravity[1], m.gravity[2] = 0.0, 0.0, -650.0
    m.car_mass, m.car_world_friction, m.car_world_restitution = 180.0, 0.3, 0.3
    m.ball_mass, m.ball_max_speed, m.ball_drag, m.ball_world_friction, m.ball_world_restitution = 30.0, 6000.0, 0.03, 0.35, 0.6
    m.jump_accel, m.jump_immediate_force = 4375.0 / 3.0, 875.0 / 3.0
    m.boost_accel_ground, m.boost_accel_air, m.boost_used_per_second = 2975.0 / 3.0, 3175.0 / 3.0, 100.0 / 3.0
    m.respawn_delay, m.bump_cooldown_time = 3.0, 0.25
    m.boost_pad_cooldown_big, m.boost_pad_cooldown_small = 10.0, 4.0
    m.car_spawn_boost_amount = 100.0 / 3.0
    m.ball_hit_extra_force_scale, m.bump_force_scale = 1.0, 1.0
    m.ball_radius = 91.25
    m.goal_base_threshold_y = 5124.25
    return m


class EngineCfg(C.Structure):
    _fields_ = [
        ("num_arenas", C.c_int32), ("team_size", C.c_int32), ("spawn_opponents", C.c_int32),
        ("tick_skip", C.c_int32), ("device", C.c_int32), ("seed", C.c_uint64), ("arena_id_base", C.c_int32),
        ("obs_kind", C.c_int32), ("obs_max_players", C.c_int32),
        ("num_reward_terms", C.c_int32), ("reward_terms", RewardTerm * RLG_MAX_REWARD_TERMS),
        ("zero_sum", C.c_int32), ("team_spirit", C.c_float), ("opponent_scale", C.c_float),
        ("no_touch_max_steps", C.c_int32), ("goal_score_terminal", C.c_int32),
        ("state_setter", C.c_int32), ("rand_ball_speed", C.c_int32), ("rand_car_speed", C.c_int32),
        ("cars_on_ground", C.c_int32),
        ("car_preset", C.c_int32), ("mutators_set", C.c_int32), ("mutators", Mutators),
    ]


def default_cfg(num_arenas: int = 256, team_size: int = 1, tick_skip: int = 8, seed: int = 123) -> EngineCfg:
    """examplemain.cpp:58-151 — the reference's canonical configuration (BASELINE cfg 1)."""
    cfg = EngineCfg()
    cfg.num_arenas = num_arenas
    cfg.team_size = team_size
    cfg.spawn_opponents = 1
    cfg.tick_skip = tick_skip
    cfg.device = 0
    cfg.seed = seed
    cfg.arena_id_base = 0
    cfg.obs_kind = RLG_OBS_DEFAULT
    cfg.obs_max_players = 3
    terms = [
        (RLG_REW_FACE_BALL, 0.1, []),
        (RLG_REW_VEL_PLAYER_TO_BALL, 0.5, []),
        (RLG_REW_VEL_BALL_TO_GOAL, 1.0, [0.0]),
        # EventReward{teamGoal 1, concede -1} * 50 (examplemain.cpp:71-76)
        (RLG_REW_EVENT, 50.0, [0, 1.0, -1.0, 0, 0, 0, 0, 0, 0, 0, 0]),
    ]
    cfg.num_reward_terms = len(terms)
    for i, (k, w, p) in enumerate(terms):
        cfg.reward_terms[i].kind = k
        cfg.reward_terms[i].weight = w
        for j, v in enumerate(p):
            cfg.reward_terms[i].params[j] = v
    cfg.zero_sum = 0
    cfg.team_spirit = 0.0
    cfg.opponent_scale = 1.0
    cfg.no_touch_max_steps = 150  # 10 s * 120 / 8 (examplemain.cpp:79)
    cfg.goal_score_terminal = 1
    cfg.state_setter = RLG_SETTER_RANDOM
    cfg.rand_ball_speed = cfg.rand_car_speed = cfg.cars_on_ground = 1
    cfg.mutators = default_mutators()
    cfg.mutators_set = 0
    return cfg


def num_players(cfg: EngineCfg) -> int:
    return cfg.team_size * (2 if cfg.spawn_opponents else 1)


def obs_size(cfg: EngineCfg) -> int:
    """51 + 19*P (DefaultOBS.cpp:20-55) or 51 + 19*2*maxPlayers (DefaultOBSPadded.cpp:3-66)."""
    if cfg.obs_kind == RLG_OBS_PADDED:
        return 51 + 19 * 2 * cfg.obs_max_players
    return 51 + 19 * num_players(cfg)


# numpy views -----------------------------------------------------------------
def _np_dtype(struct_cls):
    return np.dtype(struct_cls)


CAR_DTYPE = _np_dtype(CarState)
BALL_DTYPE = _np_dtype(BallState)
PAD_DTYPE = _np_dtype(PadState)
CONTROLS_DTYPE = _np_dtype(Controls)


def new_cars(n: int) -> np.ndarray:
    """n default CarStates (reference Car.h:17-101 defaults: on ground at z=17, boost 33.33)."""
    a = np.zeros(n, dtype=CAR_DTYPE)
    a["pos"][:, 2] = 17.0
    a["rot_forward"][:, 0] = 1.0
    a["rot_right"][:, 1] = 1.0
    a["rot_up"][:, 2] = 1.0
    a["is_on_ground"] = 1
    a["boost"] = np.float32(100.0 / 3.0)
    a["hit_tick"] = -1
    a["hit_extra_tick"] = -1
    return a


def new_balls(n: int) -> np.ndarray:
    a = np.zeros(n, dtype=BALL_DTYPE)
    a["pos"][:, 2] = 93.15
    return a


def new_pads(n: int) -> np.ndarray:
    a = np.zeros(n, dtype=PAD_DTYPE)
    a["is_active"] = 1
    return a
