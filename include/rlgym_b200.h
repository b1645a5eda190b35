/*
 * rlgym_b200.h — C ABI of the B200-native collection engine.
 *
 * This is the drop-in boundary for the reference's ThreadAgentManager/ThreadAgent
 * collection path.  Plain C: pointers, sizes and int status codes,
 * no C++ or torch types.  Every entry point cites the reference interface it
 * replaces (paths relative to /root/reference/RLGymPPO_CPP unless stated):
 *
 *   G/ = RLGymSim_CPP/src/RLGymSim_CPP/         (gym + plugin layer)
 *   R/ = RLGymSim_CPP/RocketSim/src/            (RocketSim game logic)
 *   P/ = src/                                   (PPO learner library)
 *
 * All arrays are DEVICE pointers unless the parameter name ends in _host.
 * All calls return RLG_OK (0) or a negative error; rlg_last_error() gives the
 * message (the reference throws std::runtime_error via RG_ERR_CLOSE,
 * G/Framework.h:17-22 — the C++ shim re-throws on non-zero status).
 * There is NO CPU fallback: if no CUDA device is usable every call that needs
 * one fails with RLG_ERR_CUDA.
 */
#ifndef RLGYM_B200_H
#define RLGYM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLG_OK 0
#define RLG_ERR_INVALID (-1)
#define RLG_ERR_CUDA (-2)
#define RLG_ERR_STATE (-3)

#define RLG_MAX_CARS 6   /* 3v3 */
#define RLG_NUM_PADS 34  /* R/RLConst.h:192-253 */
#define RLG_NUM_ACTIONS 90 /* G/Utils/ActionParsers/DiscreteAction.cpp:3-67 */
#define RLG_MAX_ACTIONS 96 /* widest action head / table (rlg_engine_set_action_table) */

/* ---- host-side AoS state structs (parity injection / extraction) ----------
 * Mirrors of RocketSim's CarState (R/Sim/Car/Car.h:17-115), BallState
 * (R/Sim/Ball/Ball.h), BoostPadState (R/Sim/BoostPad/BoostPad.h) and CarControls
 * (R/Sim/CarControls.h), plus the per-wheel values btVehicleRL carries between
 * ticks (R/Sim/btVehicleRL/btVehicleRL.h:9-27; SURVEY.md §7 hard part 3).
 * Every field is 4 bytes wide (bools are int32) except the two tick stamps. */
typedef struct rlg_controls {
    float throttle, steer, pitch, yaw, roll;
    int32_t jump, boost, handbrake;
} rlg_controls;

typedef struct rlg_car_state {
    float pos[3];
    float rot_forward[3], rot_right[3], rot_up[3]; /* RotMat columns */
    float vel[3];
    float ang_vel[3];
    int32_t is_on_ground;
    int32_t wheels_with_contact[4];
    int32_t has_jumped, has_double_jumped, has_flipped;
    float flip_rel_torque[3];
    float jump_time, flip_time;
    int32_t is_flipping, is_jumping;
    float air_time, air_time_since_jump;
    float boost, time_spent_boosting;
    int32_t is_supersonic;
    float supersonic_time, handbrake_val;
    int32_t is_auto_flipping;
    float auto_flip_timer, auto_flip_torque_scale;
    int32_t world_contact_has;
    float world_contact_normal[3];
    int32_t car_contact_other_id; /* car id (1-based), 0 = none */
    float car_contact_cooldown;
    int32_t is_demoed;
    float demo_respawn_timer;
    /* BallHitInfo, R/Sim/BallHitInfo/BallHitInfo.h */
    int32_t hit_valid;
    float hit_rel_pos_on_ball[3], hit_ball_pos[3], hit_extra_vel[3];
    int64_t hit_tick;         /* tickCountWhenHit; -1 == ~0ULL */
    int64_t hit_extra_tick;   /* tickCountWhenExtraImpulseApplied */
    rlg_controls last_controls;
    /* values btVehicleRL keeps from the previous tick */
    float wheel_steer_angle; /* wheels 0,1; rear wheels never steer */
    float wheel_engine_force, wheel_brake;
    float wheel_lat_friction[4], wheel_long_friction[4], wheel_extra_pushback[4];
    /* identity */
    int32_t car_id; /* 1-based, as Arena::_lastCarID assigns them */
    int32_t team;   /* 0 blue, 1 orange */
} rlg_car_state;

typedef struct rlg_ball_state {
    float pos[3];
    float vel[3];
    float ang_vel[3];
} rlg_ball_state;

typedef struct rlg_pad_state {
    int32_t is_active;
    float cooldown;
    int32_t prev_locked_car_id;
} rlg_pad_state;

/* ---- engine configuration ------------------------------------------------- */
enum rlg_obs_kind { RLG_OBS_DEFAULT = 0, RLG_OBS_PADDED = 1 };
enum rlg_car_preset { RLG_CAR_OCTANE = 0, RLG_CAR_DOMINUS = 1, RLG_CAR_PLANK = 2, RLG_CAR_BREAKOUT = 3, RLG_CAR_HYBRID = 4, RLG_CAR_MERC = 5 };
enum rlg_state_setter { RLG_SETTER_KICKOFF = 0, RLG_SETTER_RANDOM = 1, RLG_SETTER_HOST = 2 };

/* Built-in reward terms (G/Utils/RewardFunctions/CommonRewards.h). */
enum rlg_reward_kind {
    RLG_REW_EVENT = 0,              /* EventReward, params = 11 weights */
    RLG_REW_VEL_PLAYER_TO_BALL = 1, /* VelocityPlayerToBallReward */
    RLG_REW_VEL_BALL_TO_GOAL = 2,   /* VelocityBallToGoalReward, param[0] = ownGoal */
    RLG_REW_FACE_BALL = 3,          /* FaceBallReward */
    RLG_REW_VELOCITY = 4,           /* VelocityReward, param[0] = isNegative */
    RLG_REW_SAVE_BOOST = 5,         /* SaveBoostReward, param[0] = exponent (CommonRewards.h:61-70; powf: 1 ulp) */
    RLG_REW_TOUCH_BALL = 6          /* TouchBallReward, param[0] = aerialWeight (CommonRewards.h:110-124; powf: 1 ulp) */
};
#define RLG_MAX_REWARD_TERMS 8

typedef struct rlg_reward_term {
    int32_t kind;
    float weight;      /* CombinedReward weight, CombinedReward.h:36-46 */
    float params[11];
} rlg_reward_term;

/* MutatorConfig (R/Sim/MutatorConfig/MutatorConfig.h:16-72), Gym's last constructor argument (G/Gym.h:18).  Filled with the
 * soccar defaults by rlg_engine_cfg_default / rlg_mutators_default; honoured when rlg_engine_cfg.mutators_set != 0.
 * car_mass, ball_mass and ball_radius must keep their defaults (rlg_engine_create rejects other values). */
enum rlg_demo_mode { RLG_DEMO_NORMAL = 0, RLG_DEMO_ON_CONTACT = 1, RLG_DEMO_DISABLED = 2 };
typedef struct rlg_mutators {
    float gravity[3];
    float car_mass, car_world_friction, car_world_restitution;
    float ball_mass, ball_max_speed, ball_drag, ball_world_friction, ball_world_restitution;
    float jump_accel, jump_immediate_force;
    float boost_accel_ground, boost_accel_air, boost_used_per_second;
    float respawn_delay, bump_cooldown_time;
    float boost_pad_cooldown_big, boost_pad_cooldown_small;
    float car_spawn_boost_amount;
    float ball_hit_extra_force_scale, bump_force_scale;
    float ball_radius;
    int32_t unlimited_flips, unlimited_double_jumps;
    int32_t demo_mode;         /* rlg_demo_mode */
    int32_t enable_team_demos;
    float goal_base_threshold_y;
} rlg_mutators;

typedef struct rlg_engine_cfg {
    int32_t num_arenas;
    int32_t team_size;        /* Match::teamSize, G/Envs/Match.h:27-46 */
    int32_t spawn_opponents;  /* Match::spawnOpponents */
    int32_t tick_skip;        /* Gym::tickSkip, G/Gym.cpp:40-41 */
    int32_t device;           /* CUDA ordinal */
    uint64_t seed;
    int32_t arena_id_base;    /* global id of local arena 0 (multi-GPU sharding; RNG streams keyed by global id) */
    /* obs */
    int32_t obs_kind;         /* DefaultOBS / DefaultOBSPadded */
    int32_t obs_max_players;  /* DefaultOBSPadded::maxPlayers */
    /* reward graph: CombinedReward(terms) optionally wrapped in ZeroSumReward */
    int32_t num_reward_terms;
    rlg_reward_term reward_terms[RLG_MAX_REWARD_TERMS];
    int32_t zero_sum;         /* 1 = wrap in ZeroSumReward (ZeroSumReward.cpp:3-29) */
    float team_spirit, opponent_scale;
    /* terminal conditions */
    int32_t no_touch_max_steps; /* NoTouchCondition(maxSteps); <=0 disables */
    int32_t goal_score_terminal; /* GoalScoreCondition */
    /* state setter */
    int32_t state_setter;     /* KickoffState / RandomState / host-provided */
    int32_t rand_ball_speed, rand_car_speed, cars_on_ground; /* RandomState flags */
    /* car */
    int32_t car_preset;       /* Gym's CarConfig argument (G/Gym.h:18): 0 OCTANE (default), 1 DOMINUS, 2 PLANK, 3 BREAKOUT, 4 HYBRID,
                                 5 MERC — hitbox, wheel and suspension geometry of R/Sim/Car/CarConfig/CarConfig.cpp:20-88 */
    int32_t mutators_set;     /* 0: default MutatorConfig(SOCCAR); 1: use `mutators` */
    rlg_mutators mutators;
} rlg_engine_cfg;

typedef struct rlg_engine rlg_engine;

const char* rlg_last_error(void);
int rlg_abi_version(void);
size_t rlg_sizeof_car_state(void);
size_t rlg_sizeof_engine_cfg(void);

/* Fills cfg with the examplemain.cpp:58-151 configuration (1v1, DefaultOBS,
 * rewards {FaceBall .1, VelPlayerToBall .5, VelBallToGoal 1, Event{teamGoal 1, concede -1}*50},
 * NoTouch 150 + GoalScore, RandomState(true,true,true), tickSkip 8). */
void rlg_engine_cfg_default(rlg_engine_cfg* cfg);
void rlg_mutators_default(rlg_mutators* m); /* MutatorConfig(GameMode::SOCCAR) */

/* Replaces `new Gym(match, tickSkip)` x num_arenas in ThreadAgent::ThreadAgent
 * (P/private/RLGymPPO_CPP/Threading/ThreadAgent.cpp:197-206, G/Gym.cpp:40-56). */
int rlg_engine_create(const rlg_engine_cfg* cfg, rlg_engine** out);
int rlg_engine_destroy(rlg_engine* e);

/* Replaces RocketSim::Init / InitFromMem (R/RocketSim.cpp:70-212): same .cmf
 * bytes the reference reads (format: R/CollisionMeshFile/CollisionMeshFile.cpp:11-35).
 * Must be called before the first reset/step. Mesh order = body order. */
int rlg_engine_load_meshes(rlg_engine* e, const void* const* cmf_blobs_host,
                           const size_t* sizes_host, int n);

/* Gym::Reset (G/Gym.cpp:58-66) on every arena whose mask byte is non-zero
 * (NULL = all): runs the state setter + Match::EpisodeReset and writes obs. */
int rlg_engine_reset(rlg_engine* e, const uint8_t* mask_host, void* stream);

/* Car::SetState / Ball::SetState / BoostPad::SetState (R/Sim/Car/Car.cpp:23-36,
 * R/Sim/Ball/Ball.cpp:35-49) for n arenas; cars is [n * num_cars], pads [n * 34].
 * Any pointer may be NULL to leave that part untouched. Host pointers. */
int rlg_engine_set_state(rlg_engine* e, const int32_t* arena_ids_host, int n,
                         const rlg_car_state* cars_host, const rlg_ball_state* balls_host,
                         const rlg_pad_state* pads_host, const int64_t* tick_counts_host);
/* Car::GetState / Ball::GetState (R/Sim/Car/Car.cpp:10-21, Ball.cpp:27-33). */
int rlg_engine_get_state(rlg_engine* e, const int32_t* arena_ids_host, int n,
                         rlg_car_state* cars_host, rlg_ball_state* balls_host,
                         rlg_pad_state* pads_host, int64_t* tick_counts_host);

/* Arena::Step(nticks) with explicit controls (R/Sim/Arena/Arena.cpp:716-812).
 * controls is a DEVICE array [num_arenas * num_cars] in car-id order. */
int rlg_engine_tick(rlg_engine* e, const rlg_controls* controls, int nticks, void* stream);

/* Gym::Step (G/Gym.cpp:68-102) + GameInst::Step auto-reset
 * (P/public/RLGymPPO_CPP/Threading/GameInst.cpp:7-38) for every arena:
 * parse -> 1 tick -> event tracker + snapshot -> (tick_skip-1) ticks ->
 * obs/reward/done -> reset finished arenas.  action_idx: DEVICE int32 [num_arenas * num_cars]
 * in player order (see rlg_engine_player_order). */
int rlg_engine_step(rlg_engine* e, const int32_t* action_idx, void* stream);

/* Device pointers to the step outputs; valid until the next step/reset.
 * obs [A*P, obs_size] f32 (post-reset obs for finished arenas, as GameInst::Step
 * stores), reward [A*P] f32, done [A] u8, next_obs_terminal is not kept: the
 * reference never uses it (P/private/RLGymPPO_CPP/Util/TorchFuncs.cpp:24,36). */
int rlg_engine_outputs(rlg_engine* e, float** obs, float** reward, uint8_t** done);
int rlg_engine_obs_size(const rlg_engine* e);
int rlg_engine_num_players(const rlg_engine* e); /* per arena */
/* players[i] -> car id; the reference's order is unordered_set<Car*> iteration
 * order (R/Sim/Arena/Arena.h:35), i.e. descending car id for the small sets used. */
int rlg_engine_player_order(const rlg_engine* e, int32_t* car_ids_host);

/* DiscreteAction table, 90x8 f32 (G/Utils/ActionParsers/DiscreteAction.cpp:3-67). */
int rlg_action_table(float* table_host);

/* Gym::Reset WITHOUT the state setter: Match::EpisodeReset + obs for whatever state the masked
 * arenas currently hold (after rlg_engine_set_state). This is how a host/custom StateSetter
 * plugs in (StateSetter::ResetState(Arena*), G/Utils/StateSetters/StateSetter.h:9) and how the
 * parity tests obtain obs for injected states (G/Envs/Match.cpp:4-23). */
int rlg_engine_reset_current(rlg_engine* e, const uint8_t* mask_host, void* stream);

/* Same, writing the masked arenas' obs rows into a caller DEVICE buffer [A*P, obs] (e.g. a trajectory-ring slot). */
int rlg_engine_reset_current_to(rlg_engine* e, const uint8_t* mask_host, float* obs_out, void* stream);

/* Match::BuildObservations / IsDone / GetRewards (G/Envs/Match.cpp:12-38) on the CURRENT arena states, i.e.
 * Gym::Step (G/Gym.cpp:84-93) minus the physics ticks and the event tracker: prevActions := table[action_idx]
 * (zeroed for demoed players), GameState::UpdateFromArena, obs, done, rewards -> outputs. action_idx: DEVICE [A*P]. */
int rlg_engine_eval_gym(rlg_engine* e, const int32_t* action_idx, void* stream);

/* Gym::Step without GameInst's auto-reset: finished arenas keep their terminal state and obs. */
int rlg_engine_step_noreset(rlg_engine* e, const int32_t* action_idx, void* stream);

/* Sets players[i] -> car id (a permutation of 1..P). Default: descending id, which is what the
 * reference's unordered_set<Car*> yields for 1v1; tests read the order back from the oracle. */
int rlg_engine_set_player_order(rlg_engine* e, const int32_t* car_ids_host);

int rlg_engine_num_arenas(const rlg_engine* e);
int rlg_engine_device(const rlg_engine* e); /* CUDA ordinal the engine lives on */
int rlg_engine_arena_id_base(const rlg_engine* e); /* rlg_engine_cfg.arena_id_base */
size_t rlg_engine_state_bytes_per_arena(const rlg_engine* e); /* S(P) of SURVEY.md §8d as laid out here */
/* D2H copy of the current outputs (any pointer may be NULL); synchronises the engine stream. */
int rlg_engine_read_outputs(rlg_engine* e, float* obs_host, float* reward_host, uint8_t* done_host);
void* rlg_engine_stream(rlg_engine* e); /* the engine's own cudaStream_t */

/* Host-buffer convenience used by the C++ shim / bench e2e: copies action_idx
 * from host, steps, and copies obs/reward/done back into host buffers. */
int rlg_engine_step_host(rlg_engine* e, const int32_t* action_idx_host,
                         float* obs_host, float* reward_host, uint8_t* done_host);

/* Zero-copy variant of rlg_engine_step_host: the engine owns page-locked host buffers (action_idx [A*P] i32 in;
 * obs [A*P, obs] f32, reward [A*P] f32, done [A] u8 out — the StepResult of G/Gym.h:24-29 for every arena); the
 * caller writes action indices into *action_idx, calls rlg_engine_step_pinned and reads the results in place.
 * The pointers stay valid for the engine's lifetime; the outputs are overwritten by the next host-buffer step. */
int rlg_engine_host_buffers(rlg_engine* e, int32_t** action_idx, float** obs, float** reward, uint8_t** done);
int rlg_engine_step_pinned(rlg_engine* e, int want_obs);

/* Synchronous copies ordered after everything queued on the engine's stream (test / host-plugin plumbing). */
int rlg_engine_copy_to_host(rlg_engine* e, void* dst_host, const void* src_dev, size_t bytes);
int rlg_engine_copy_to_device(rlg_engine* e, void* dst_dev, const void* src_host, size_t bytes);

/* GameInst's reward metrics (P/public/RLGymPPO_CPP/Threading/GameInst.cpp:13-31: avgStepRew.Add(sum of the players'
 * rewards, playerAmount); curEpRew += that sum / playerAmount; on done avgEpRew += curEpRew) accumulated on the device
 * per arena by every fused Gym::Step, and summed over the arenas in arena order like
 * ThreadAgentManager::GetMetrics (P/private/RLGymPPO_CPP/Threading/ThreadAgentManager.cpp:82-92).  The averages are
 * NaN while their count is 0 (AvgTracker::Get, P/public/RLGymPPO_CPP/Util/AvgTracker.h:14-20). */
typedef struct rlg_metrics_host {
    float avg_step_reward;      /* report["Average Step Reward"] */
    float avg_episode_reward;   /* report["Average Episode Reward"] */
    float step_reward_total;
    float episode_reward_total;
    uint64_t step_reward_count; /* player-steps */
    uint64_t episode_count;     /* finished episodes */
    uint64_t total_steps;       /* GameInst::totalSteps summed over the arenas (never reset) */
} rlg_metrics_host;
int rlg_engine_metrics(rlg_engine* e, rlg_metrics_host* out);   /* synchronises the engine stream */
int rlg_engine_reset_metrics(rlg_engine* e);                    /* GameInst::ResetMetrics for every arena */

/* GameState::scoreLine of every arena (G/Utils/Gamestates/GameState.cpp:100-101: incremented by the step's snapshot when the
 * ball is behind a goal line, index 0 = ball at y > 0 = blue scored; zeroed by a reset): out_host is [A][2] int32.
 * What SkillTracker's goal test (Math::IsBallScored on stepResult.state, SkillTracker.cpp:132-134) reads. */
int rlg_engine_score_lines(rlg_engine* e, int32_t* out_host);

/* Number of kernel launches issued by this engine so far (bench "gpu_launches"). */
uint64_t rlg_engine_launch_count(const rlg_engine* e);
/* Wait for all work queued on the engine's stream. */
int rlg_engine_sync(rlg_engine* e);


/* Gym::Step with caller-provided DEVICE output buffers (obs [A*P,obs], reward [A*P], done [A]); the engine's own
 * output buffers (rlg_engine_outputs) are left untouched. This is how the collector appends straight into its
 * trajectory ring (replaces GameTrajectory::AppendSingleStep, P/private/RLGymPPO_CPP/Threading/GameTrajectory.cpp:54-79). */
int rlg_engine_step_to(rlg_engine* e, const int32_t* action_idx, float* obs_out, float* reward_out, uint8_t* done_out,
                       void* stream);
/* Per-block completion flags of the LAST rlg_engine_step / _step_to launch, for a consumer kernel that wants to start on the
 * outputs of the arena blocks that are done while the slowest blocks of the step are still running (the collector's inference
 * kernel, launched as a programmatic dependent of the step): block b of the step's grid holds arenas
 * [b * arenas_per_block, (b + 1) * arenas_per_block) and stores flags_dev[b] = seq (device scope, after its obs / reward / done /
 * state stores) when it is finished.  flags_dev is NULL when the engine does not publish flags (per-group barrier modes).
 * No reference counterpart: the reference's agents wait for Gym::Step to return (ThreadAgent.cpp:100-140). */
int rlg_engine_step_ready(rlg_engine* e, const uint32_t** flags_dev, uint32_t* seq, int* arenas_per_block);
/* A user ActionParser (G/Utils/ActionParsers/ActionParser.h:11-14) whose ParseActions maps an action index to an Action independently
 * of the game state: its table, n_actions rows of 8 floats (throttle, steer, pitch, yaw, roll, jump, boost, handbrake — Action.h:5-9),
 * replaces the DiscreteAction table of the fused step; 1 <= n_actions <= RLG_MAX_ACTIONS.  Call it before rlg_collector_create: the
 * policy head gets rlg_engine_num_actions(e) outputs (ActionParser::GetActionAmount). */
int rlg_engine_set_action_table(rlg_engine* e, const float* table_host, int n_actions);
int rlg_engine_num_actions(const rlg_engine* e);
/* rlg_engine_step_to launched as the programmatic dependent of the kernel before it on `stream` (the inference that writes action_idx):
 * every block of the step waits until tile_flags_dev[i] >= seq for the tiles i of rows_per_tile consecutive action rows that cover its
 * arenas, instead of for the whole producer kernel.  tile_flags_dev == NULL: exactly rlg_engine_step_to. */
int rlg_engine_step_to_after(rlg_engine* e, const int32_t* action_idx, float* obs_out, float* reward_out, uint8_t* done_out, void* stream,
                             const uint32_t* tile_flags_dev, uint32_t seq, int rows_per_tile);

/* ---- host-plugin path: user-defined OBSBuilder / RewardFunction / TerminalCondition / StepCallback -------------------------------
 * (G/Utils/OBSBuilders/OBSBuilder.h:10-15, RewardFunctions/RewardFunction.h:9-35, TerminalConditions/TerminalCondition.h:7-8,
 * P/public/RLGymPPO_CPP/Threading/GameInst.h:7).  Gym::Step is split where the reference takes its GameState (G/Gym.cpp:84-93):
 * step_begin = ParseActions + the first tick + GameEventTracker::Update + GameState::UpdateFromArena (+ the fused built-in obs /
 * reward / done into the given buffers, NULL = the engine's own); export_gamestates = that GameState for the host;
 * step_end = the other tickSkip - 1 ticks.  No auto-reset and no reward metrics: the host decides `done` and keeps GameInst's
 * trackers.  Arrays: cars / players are [n * P] in PLAYER order (GameState::players), balls / gym [n]; ids NULL = every arena. */
typedef struct rlg_gym_player {   /* PlayerData minus its CarState (G/Utils/Gamestates/PlayerData.h:7-38) */
    int32_t match_goals, match_saves, match_assists, match_shots, match_shot_passes, match_bumps, match_demos, boost_pickups;
    int32_t ball_touched_step, ball_touched_tick;
    float prev_action[8];         /* Match::prevActions row */
} rlg_gym_player;
typedef struct rlg_gym_state {    /* GameState minus players / ball (G/Utils/Gamestates/GameState.h:19-57) */
    int64_t tick_count;           /* Arena::tickCount == GameState::lastTickCount after the update */
    int32_t score_line[2], last_touch_car_id;
    int32_t steps_since_touch;    /* NoTouchCondition's counter (informational) */
    int32_t pad_active[RLG_NUM_PADS];  /* GameState::boostPads order = CommonValues::BOOST_LOCATIONS */
    float pad_cooldown[RLG_NUM_PADS];  /* GameState::boostPadTimers */
} rlg_gym_state;
size_t rlg_sizeof_gym_state(void);
size_t rlg_sizeof_gym_player(void);
int rlg_engine_step_begin(rlg_engine* e, const int32_t* action_idx, float* obs_out, float* reward_out, uint8_t* done_out, void* stream);
int rlg_engine_step_end(rlg_engine* e, const int32_t* action_idx, void* stream);
int rlg_engine_export_gamestates(rlg_engine* e, const int32_t* arena_ids_host, int n, rlg_car_state* cars_host, rlg_ball_state* balls_host,
                                 rlg_gym_state* gym_host, rlg_gym_player* players_host);
/* The same in two halves so that step_end can run on the GPU while the copy is in flight and the host plugins work. */
int rlg_engine_export_gamestates_async(rlg_engine* e, const int32_t* arena_ids_host, int n, rlg_car_state* cars_host, rlg_ball_state* balls_host,
                                       rlg_gym_state* gym_host, rlg_gym_player* players_host);
int rlg_engine_export_wait(rlg_engine* e);
/* Gym::Reset (state setter + EpisodeReset) with the obs rows written to a caller DEVICE buffer [A*P, obs]. */
int rlg_engine_reset_to(rlg_engine* e, const uint8_t* mask_host, float* obs_out, void* stream);
/* Page-locked host memory for the export / upload buffers. */
void* rlg_host_alloc(size_t bytes);
void rlg_host_free(void* p);
/* Device memory on the engine's GPU for callers without a CUDA runtime of their own (the C++ shim's trajectory tensors). */
void* rlg_device_alloc(rlg_engine* e, size_t bytes);
void rlg_device_free(rlg_engine* e, void* p);
/* Lets a host callback (step / reset hook) report why it failed through rlg_last_error(). */
void rlg_set_last_error(const char* msg);

/* ---- collector: device-resident ThreadAgent loop --------------------------------------------------------------------
 * Replaces ThreadAgent::_RunFunc (P/private/RLGymPPO_CPP/Threading/ThreadAgent.cpp:24-195: infer -> step -> append),
 * DiscretePolicy::GetAction (P/private/RLGymPPO_CPP/PPO/DiscretePolicy.cpp:44-62), ValueEstimator::Forward
 * (PPO/ValueEstimator.cpp:6-27), ThreadAgentManager::CollectTimesteps' truncation marking + concatenation
 * (Threading/ThreadAgentManager.cpp:16-80), TorchFuncs::ComputeGAE (Util/TorchFuncs.cpp:5-52) and the layout of
 * ExperienceBuffer::SubmitExperience rows (PPO/ExperienceBuffer.cpp:12-70).
 *
 * Both networks are Linear+ReLU stacks with a final Linear (DiscretePolicy.cpp:7-30, ValueEstimator.cpp:6-27); the
 * forward runs on the 5th-gen tensor cores (tcgen05, TF32 inputs, FP32 accumulate in TMEM) fused with bias+ReLU,
 * softmax(logits / temperature), clamp(ACTION_MIN_PROB=1e-11, 1), multinomial sampling and log-prob. */
/* ---- PPO minibatch update: the dense contractions on the tensor cores (csrc/gemm.cu) ----------------------------------
 * C[M, N] (+)= A[M, K] . B[N, K]^T (+ bias[N]) (ReLU): row-major fp32 device matrices, TF32 tcgen05.mma with FP32
 * accumulation.  Replaces the three cuBLAS GEMMs per torch::nn::Linear that autograd runs for PPOLearner::Learn
 * (P/private/RLGymPPO_CPP/PPO/PPOLearner.cpp:125-290): forward (A = X, B = W), input gradient (A = dY, B = W^T) and
 * weight gradient (A = dY^T, B = X^T, split over K = rows with RLG_GEMM_ATOMIC).  K, lda, ldb multiples of 4 floats,
 * A and B 16-byte aligned; bias may be NULL; split_k >= 1 (> 1 needs RLG_GEMM_ATOMIC: C must hold the sum's start). */
#define RLG_GEMM_RELU 1         /* C = max(., 0) */
#define RLG_GEMM_ACCUMULATE 2   /* C += (non-atomic read-modify-write) */
#define RLG_GEMM_ATOMIC 4       /* C += with atomic adds (split-K) */
#define RLG_GEMM_SCALAR_STORE 8 /* internal: unaligned C */
int rlg_gemm_tf32(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias,
                  int flags, int split_k, void* stream);
/* The same with two fused epilogue extras (either may be NULL): mask [M, N] (ld ldm) zeroes C where mask <= 0 — the ReLU
 * backward of the input-gradient GEMM, mask = the layer's forward output — and Ct [N, M] (ld ldct) receives the
 * transposed result too, which is the K-major operand the weight-gradient GEMM of the neighbouring layer needs. */
int rlg_gemm_tf32_fused(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias,
                        int flags, int split_k, const float* mask, int ldm, float* Ct, int ldct, void* stream);

#define RLG_MAX_HIDDEN_LAYERS 4
typedef struct rlg_collector_cfg {
    int32_t num_hidden;                          /* layerSizes.size() (LearnerConfig.h policyLayerSizes / criticLayerSizes) */
    int32_t policy_hidden[RLG_MAX_HIDDEN_LAYERS]; /* multiples of 32, <= 256 */
    int32_t critic_hidden[RLG_MAX_HIDDEN_LAYERS];
    int32_t max_steps;                           /* trajectory ring capacity T (env-steps per collect call) */
    uint64_t seed;                               /* sampling RNG (counter based: seed, global row id, step counter) */
    float temperature;                           /* DiscretePolicy::temperature */
    int32_t deterministic;                       /* ThreadAgentManager::deterministic -> argmax, logprob 0 */
} rlg_collector_cfg;

typedef struct rlg_collector rlg_collector;

/* T-major device views of the last collect: row index n = arena * P + player (player order).
 * obs [T+1][N][obs] (slot t = what the policy saw at step t; slot T = nextStates of the last step),
 * action [T][N] i32, logprob [T][N], reward [T][N], done [T][A] u8, value [T+1][N] (slot T = bootstrap),
 * advantage/value_target/ret [T][N] (after rlg_collector_gae). */
typedef struct rlg_traj_view {
    int32_t T, N, A, P, obs_size;
    const float* obs; const int32_t* action; const float* logprob; const float* reward; const uint8_t* done;
    const float* value; const float* advantage; const float* value_target; const float* ret;
} rlg_traj_view;

int rlg_collector_create(rlg_engine* e, const rlg_collector_cfg* cfg, rlg_collector** out);
int rlg_collector_destroy(rlg_collector* c);
/* net: 0 = policy, 1 = critic. layer in [0, num_hidden]. W_host is torch's nn.Linear weight [out, in] row-major,
 * b_host [out]. (Same tensors PPOLearner keeps in policy->parameters(), P/private/RLGymPPO_CPP/PPO/PPOLearner.cpp.) */
int rlg_collector_set_layer(rlg_collector* c, int net, int layer, const float* W_host, const float* b_host, int out_dim, int in_dim);
/* DiscretePolicy::GetAction + ValueEstimator::Forward on obs_dev [n_rows, obs_size]; any output may be NULL.
 * counter selects the RNG stream position (the collector uses its env-step counter). */
int rlg_collector_infer(rlg_collector* c, const float* obs_dev, int n_rows, uint64_t counter, int32_t* action_dev,
                        float* logprob_dev, float* value_dev, void* stream);
/* n_steps (<= max_steps) iterations of: infer(obs_t) -> rlg_engine_step_to(ring slot t+1) ; then the bootstrap value
 * pass on slot T. The first call uses the engine's current obs (rlg_engine_reset must have run). */
int rlg_collector_collect(rlg_collector* c, int n_steps, void* stream);
/* TorchFuncs::ComputeGAE over the reference's concatenation order (player-major: each row n contributes its T steps
 * back to back; the last step of every row is marked truncated unless done; the value after a row's last step is the
 * NEXT row's first value — the reference's seam quirk — and the bootstrap value for the last row). */
int rlg_collector_gae(rlg_collector* c, float gamma, float lambda, float return_std, float clip_range, void* stream);
int rlg_collector_view(rlg_collector* c, rlg_traj_view* out);
/* ExperienceBuffer::SubmitExperience row layout: writes the last collect in the reference's concatenated order
 * (row i = n * T + t) into caller DEVICE buffers (any may be NULL): states [N*T, obs], actions i64 [N*T],
 * log_probs, rewards, next_states [N*T, obs], dones f32, truncateds f32, value_targets, advantages. */
int rlg_collector_export(rlg_collector* c, float* states, int64_t* actions, float* log_probs, float* rewards,
                         float* next_states, float* dones, float* truncateds, float* value_targets, float* advantages,
                         void* stream);
/* Host StateSetter support (rlg_engine_cfg.state_setter == RLG_SETTER_HOST): after every env-step of a collect the
 * collector synchronises, reads the done flags and calls hook(user, finished_arena_ids, n, obs_out) where obs_out is
 * the DEVICE obs slot the post-reset observations must be written to (rlg_engine_set_state +
 * rlg_engine_reset_current_to). Replaces the StateSetter::ResetState(Arena*) call in Match::ResetState
 * (G/Envs/Match.cpp:54-70) reached from GameInst::Step's auto-reset (GameInst.cpp:20-24). */
typedef void (*rlg_reset_hook)(void* user, const int32_t* arena_ids_host, int n, float* obs_out);
int rlg_collector_set_reset_hook(rlg_collector* c, rlg_reset_hook hook, void* user);
/* Host-plugin collection: with a step hook installed every env-step of a collect runs as
 *   infer -> rlg_engine_step_begin (built-in outputs into the ring slot) -> hook(user, t, actions, obs_next, reward, done)
 * and the hook finishes the step: it exports the GameStates, launches rlg_engine_step_end, runs the user's plugins, uploads the rows
 * they produced into the DEVICE slots it was given (obs_next [A*P, obs], reward [A*P], done [A] u8) and re-sets finished arenas
 * (writing their post-reset obs into obs_next).  Replaces GameInst::Step's body for that configuration (GameInst.cpp:7-38). */
typedef int (*rlg_step_hook)(void* user, int t, const int32_t* actions_dev, float* obs_next_dev, float* reward_dev, uint8_t* done_dev);
int rlg_collector_set_step_hook(rlg_collector* c, rlg_step_hook hook, void* user);
/* A trajectory collected elsewhere (HOST arrays in the ring's T-major layout: obs [T+1][N][obs], action [T][N] i32, logprob / reward
 * [T][N], done [T][A] u8, value [T+1][N]) becomes the collector's last collect, so that rlg_collector_gae / _return_stats / _export
 * and rlg_ppo_submit_collector work on it (e.g. timesteps gathered by CPU RLGymSim gyms next to the device pool). */
int rlg_collector_load_external(rlg_collector* c, int n_steps, const float* obs_host, const int32_t* action_host, const float* logprob_host,
                                const float* reward_host, const uint8_t* done_host, const float* value_host);
uint64_t rlg_collector_launch_count(const rlg_collector* c);
/* Per-kernel CUDA-event timing of the LAST collect on its launching stream (bench roofline): summed durations and launch
 * counts of the fused Gym::Step kernel and of the MLP inference kernel. Replaces ThreadAgent::Times
 * (P/private/RLGymPPO_CPP/Threading/ThreadAgent.h "envStepTime"/"policyInferTime", ThreadAgentManager.cpp:82-117). */
int rlg_collector_enable_timing(rlg_collector* c, int on);
int rlg_collector_kernel_times(rlg_collector* c, double* step_ms, int32_t* step_launches, double* infer_ms, int32_t* infer_launches);
/* rlg_collector_set_layer from DEVICE memory (W_dev [out, in] row-major, leading dimension ldw; b_dev [out]), packed by a kernel on
 * `stream` (NULL = the engine's stream) — no host round trip. */
int rlg_collector_set_layer_device(rlg_collector* c, int net, int layer, const float* W_dev, int ldw, const float* b_dev, int out_dim, int in_dim,
                                   void* stream);
/* Learner::AddNewExperience's report values (P/public/RLGymPPO_CPP/Learner.cpp:660-682) of the last collect + GAE, reduced on the
 * device: out3_host = { mean |returns|, mean |advantages|, mean |value targets| } and the first n_first returns in the
 * reference's concatenation order (what WelfordRunningStat::Increment receives, maxReturnsPerStatsInc). Synchronises. */
int rlg_collector_return_stats(rlg_collector* c, double* out3_host, float* first_returns_host, int n_first, void* stream);

/* ---- PPO learner on the device (csrc/ppo.cu) ---------------------------------------------------------------------------------
 * Replaces PPOLearner (P/private/RLGymPPO_CPP/PPO/PPOLearner.cpp:17-349, 504-517) and ExperienceBuffer
 * (PPO/ExperienceBuffer.cpp:12-121): the networks' parameters, gradients and Adam moments are ONE flat device vector each
 * ([policy | critic], weights padded to multiples of 4 floats per dimension), the experience FIFO is a device ring, and
 * Learn is hand-written kernels + the tcgen05 GEMM above — no autograd, no torch. */
typedef struct rlg_ppo_cfg {
    int32_t device;
    int32_t obs_size, num_actions;
    int32_t num_hidden;                             /* layerSizes.size() */
    int32_t policy_hidden[RLG_MAX_HIDDEN_LAYERS];   /* PPOLearnerConfig::policyLayerSizes (multiples of 4) */
    int32_t critic_hidden[RLG_MAX_HIDDEN_LAYERS];   /* PPOLearnerConfig::criticLayerSizes */
    int64_t batch_size, mini_batch_size;            /* per replica; mini_batch_size 0 = batch_size (PPOLearner.cpp:19-23) */
    int32_t epochs;
    float policy_lr, critic_lr, ent_coef, clip_range, temperature;
    int64_t exp_buffer_size;                        /* ExperienceBuffer maxSize, per replica */
    uint64_t seed;                                  /* shuffle stream */
    int32_t world;                                  /* data-parallel replicas (gradients = all-reduce SUM / world) */
} rlg_ppo_cfg;
typedef struct rlg_ppo_report {   /* the report keys PPOLearner::Learn writes (PPOLearner.cpp:296-347) */
    double entropy, kl, ratio, value_loss, clip_fraction;      /* means over the minibatches */
    double policy_update_magnitude, critic_update_magnitude;   /* |params before - after| */
    int64_t batches, minibatches, cumulative_model_updates;
    double device_ms;                                          /* CUDA-event time of the call on its stream */
} rlg_ppo_report;
typedef struct rlg_ppo rlg_ppo;
/* In-place SUM all-reduce of count floats at grads_dev over the replicas, enqueued on `stream` (NCCL in the Python host). */
typedef void (*rlg_allreduce_hook)(void* user, float* grads_dev, int64_t count, void* stream);

int rlg_ppo_create(const rlg_ppo_cfg* cfg, rlg_ppo** out);
int rlg_ppo_destroy(rlg_ppo* p);
/* torch::nn::Linear default initialisation of every layer (DiscretePolicy.cpp:13-27, ValueEstimator.cpp:10-24). */
int rlg_ppo_init_weights(rlg_ppo* p, uint64_t seed);
/* which: 0 parameters, 1 gradients, 2 Adam exp_avg, 3 Adam exp_avg_sq; net: 0 policy, 1 critic; W_host [out, in], b_host [out]. */
int rlg_ppo_set_layer(rlg_ppo* p, int which, int net, int layer, const float* W_host, const float* b_host, int out_dim, int in_dim);
int rlg_ppo_get_layer(rlg_ppo* p, int which, int net, int layer, float* W_host, float* b_host, int out_dim, int in_dim);
int rlg_ppo_adam_steps(rlg_ppo* p, int64_t* policy_steps, int64_t* critic_steps, int set);
int rlg_ppo_flat(rlg_ppo* p, int which, float** dev, int64_t* count, int64_t* policy_count);
int rlg_ppo_set_lr(rlg_ppo* p, float policy_lr, float critic_lr); /* PPOLearner::UpdateLearningRates */
int rlg_ppo_set_allreduce_hook(rlg_ppo* p, rlg_allreduce_hook hook, void* user, int world);
/* ExperienceBuffer::SubmitExperience: n rows of DEVICE data (states [n, obs], actions i64, log_probs, value targets, advantages). */
int rlg_ppo_submit(rlg_ppo* p, const float* states, const int64_t* actions, const float* log_probs, const float* value_targets,
                   const float* advantages, int64_t n, void* stream);
int rlg_ppo_submit_collector(rlg_ppo* p, rlg_collector* c, void* stream); /* the collector's last collect (after rlg_collector_gae) */
int64_t rlg_ppo_buffer_size(const rlg_ppo* p);
int rlg_ppo_buffer_read(rlg_ppo* p, float* states_host, int64_t* actions_host, float* log_probs_host, float* value_targets_host,
                        float* advantages_host);   /* FIFO order, oldest first */
int rlg_ppo_peek_shuffle(rlg_ppo* p, int32_t* perm_host, uint64_t counter); /* the permutation epoch `counter` uses */
uint64_t rlg_ppo_shuffle_counter(const rlg_ppo* p);
/* PPOLearner::Learn on the buffer's current contents; report may be NULL (then the call does not synchronise). */
int rlg_ppo_learn(rlg_ppo* p, rlg_ppo_report* report, void* stream);
/* New weights to the agents (ThreadAgentManager::SetNewPolicy + the critic), device to device. */
int rlg_ppo_push_weights(rlg_ppo* p, rlg_collector* c, void* stream);
void* rlg_ppo_stream(rlg_ppo* p);
uint64_t rlg_ppo_launch_count(const rlg_ppo* p);
int64_t rlg_ppo_model_updates(const rlg_ppo* p);

#ifdef __cplusplus
}
#endif
#endif /* RLGYM_B200_H */
