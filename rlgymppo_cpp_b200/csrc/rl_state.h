// rl_state.h — per-arena simulation state and the constants of the reference's soccar mode.
//
// One ArenaS holds everything a RocketSim Arena + RLGymSim Gym keep between ticks
// (reference Arena.h, Car.h:17-115, Ball.h, BoostPad.h, GameEventTracker.h:60-70,
// Gym.h:8-15).  In HBM the engine stores the arenas word-transposed
// (word w of arena a lives at buf[w * num_arenas + a]) so that the 32 threads of a warp,
// one arena each, read and write fully coalesced 128-byte lines.  Every member is
// therefore 4 bytes wide (int64 tick stamps are split hi/lo).
#pragma once
#include "rl_math.h"

namespace rl {

constexpr int kMaxCars = 6;
constexpr int kNumPads = 34;
constexpr int kNumPadsBig = 6;
constexpr float UU2BT = 1.f / 50.f;
constexpr float BT2UU = 50.f;
constexpr float kTickTime = 1.f / 120.f;

// ---- RLConst.h ---------------------------------------------------------------------
namespace C {
constexpr float GRAVITY_Z = -650.f;
constexpr float ARENA_EXTENT_X = 4096, ARENA_EXTENT_Y = 5120, ARENA_HEIGHT = 2048;
constexpr float CAR_MASS = 180.f, BALL_MASS = CAR_MASS / 6.f;
constexpr float CARBALL_FRICTION = 2.0f, CARBALL_RESTITUTION = 0.0f;
constexpr float CARWORLD_FRICTION = 0.3f, CARWORLD_RESTITUTION = 0.3f;
constexpr float CARCAR_FRICTION = 0.09f, CARCAR_RESTITUTION = 0.1f;
constexpr float BALL_REST_Z = 93.15f, BALL_MAX_ANG_SPEED = 6.f, BALL_DRAG = 0.03f;
constexpr float BALL_FRICTION = 0.35f, BALL_RESTITUTION = 0.6f;
constexpr float WORLD_FRICTION = 0.6f, WORLD_RESTITUTION = 0.3f;  // Arena.cpp:503-505
constexpr float CAR_MAX_SPEED = 2300.f, BALL_MAX_SPEED = 6000.f;
constexpr float BOOST_MAX = 100.f, BOOST_USED_PER_SECOND = BOOST_MAX / 3, BOOST_MIN_TIME = 0.1f;
constexpr float BOOST_ACCEL_GROUND = 2975 / 3.f, BOOST_ACCEL_AIR = 3175 / 3.f, BOOST_SPAWN_AMOUNT = BOOST_MAX / 3;
constexpr float CAR_MAX_ANG_SPEED = 5.5f;
constexpr float SUPERSONIC_START_SPEED = 2200.f, SUPERSONIC_MAINTAIN_MIN_SPEED = SUPERSONIC_START_SPEED - 100.f;
constexpr float SUPERSONIC_MAINTAIN_MAX_TIME = 1.f;
constexpr float POWERSLIDE_RISE_RATE = 5, POWERSLIDE_FALL_RATE = 2;
constexpr float THROTTLE_TORQUE_AMOUNT = CAR_MASS * 400.f;
constexpr float BRAKE_TORQUE_AMOUNT = CAR_MASS * (14.25f + (1.f / 3.f));
constexpr float STOPPING_FORWARD_VEL = 25.f, COASTING_BRAKE_FACTOR = 0.15f;
constexpr float BRAKING_NO_THROTTLE_SPEED_THRESH = 0.01f, THROTTLE_DEADZONE = 0.001f;
constexpr float THROTTLE_AIR_ACCEL = 200 / 3.f;
constexpr float JUMP_ACCEL = 4375.f / 3.f, JUMP_IMMEDIATE_FORCE = 875.f / 3.f, JUMP_MIN_TIME = 0.025f;
constexpr float JUMP_RESET_TIME_PAD = (1 / 40.f), JUMP_MAX_TIME = 0.2f, DOUBLEJUMP_MAX_DELAY = 1.25f;
constexpr float FLIP_Z_DAMP_120 = 0.35f, FLIP_Z_DAMP_START = 0.15f, FLIP_Z_DAMP_END = 0.21f;
constexpr float FLIP_TORQUE_TIME = 0.65f, FLIP_TORQUE_MIN_TIME = 0.41f, FLIP_PITCHLOCK_TIME = 1.f;
constexpr float FLIP_PITCHLOCK_EXTRA_TIME = 0.3f, FLIP_INITIAL_VEL_SCALE = 500.f;
constexpr float FLIP_TORQUE_X = 260.f, FLIP_TORQUE_Y = 224.f;
constexpr float FLIP_FORWARD_IMPULSE_MAX_SPEED_SCALE = 1.f, FLIP_SIDE_IMPULSE_MAX_SPEED_SCALE = 1.9f;
constexpr float FLIP_BACKWARD_IMPULSE_MAX_SPEED_SCALE = 2.5f, FLIP_BACKWARD_IMPULSE_SCALE_X = 16.f / 15.f;
constexpr float BALL_RADIUS = 91.25f;
constexpr float GOAL_THRESHOLD_Y = 5124.25f;
constexpr float CAR_TORQUE_SCALE = (float)(2 * 3.14159265358979323846 / (1 << 16) * 1000);
constexpr float CAR_AUTOFLIP_IMPULSE = 200, CAR_AUTOFLIP_TORQUE = 50, CAR_AUTOFLIP_TIME = 0.4f;
constexpr float CAR_AUTOFLIP_NORMZ_THRESH = (float)0.70710678118654752440, CAR_AUTOFLIP_ROLL_THRESH = 2.8f;
constexpr float CAR_AUTOROLL_FORCE = 100, CAR_AUTOROLL_TORQUE = 80;
constexpr float BALL_CAR_EXTRA_IMPULSE_Z_SCALE = 0.35f, BALL_CAR_EXTRA_IMPULSE_FORWARD_SCALE = 0.65f;
constexpr float BALL_CAR_EXTRA_IMPULSE_MAXDELTAVEL_UU = 4600.f;
constexpr float CAR_SPAWN_REST_Z = 17.f, CAR_RESPAWN_Z = 36.f;
constexpr float BUMP_COOLDOWN_TIME = 0.25f, BUMP_MIN_FORWARD_DIST = 64.5f, DEMO_RESPAWN_TIME = 3.f;
constexpr float SUSPENSION_FORCE_SCALE_FRONT = 36.f - (1.f / 4.f);
constexpr float SUSPENSION_FORCE_SCALE_BACK = 54.f + (1.f / 4.f) + (1.5f / 100.f);
constexpr float SUSPENSION_STIFFNESS = 500.f, WHEELS_DAMPING_COMPRESSION = 25.f, WHEELS_DAMPING_RELAXATION = 40.f;
constexpr float MAX_SUSPENSION_TRAVEL = 12.f, SUSPENSION_SUBTRACTION = 0.05f;
constexpr float AIR_TORQUE_P = 130, AIR_TORQUE_Y = 95, AIR_TORQUE_R = 400;
constexpr float AIR_DAMP_P = 30, AIR_DAMP_Y = 20, AIR_DAMP_R = 50;
constexpr float PAD_CYL_HEIGHT = 95, PAD_CYL_RAD_BIG = 208, PAD_CYL_RAD_SMALL = 144;
constexpr float PAD_BOX_HEIGHT = 64, PAD_BOX_RAD_BIG = 160, PAD_BOX_RAD_SMALL = 120;
constexpr float PAD_COOLDOWN_BIG = 10, PAD_COOLDOWN_SMALL = 4, PAD_BOOST_BIG = 100, PAD_BOOST_SMALL = 12;
// car presets (R/Sim/Car/CarConfig/CarConfig.cpp:20-88): OCTANE, DOMINUS, PLANK, BREAKOUT, HYBRID, MERC
constexpr int kNumCarPresets = 6;
struct CarPreset { float hitbox[3], hitboxOff[3], wheelRFront, wheelRBack, susRestFront, susRestBack, wheelF[3], wheelB[3]; };
RL_HDI CarPreset car_preset(int i) {
    const CarPreset T[kNumCarPresets] = {
        {{120.507f, 86.6994f, 38.6591f}, {(float)13.87566, 0.f, 20.755f}, 12.50f, 15.00f, 38.755f, 37.055f, {51.25f, 25.90f, 20.755f}, {-33.75f, 29.50f, 20.755f}},
        {{130.427f, 85.7799f, 33.8f}, {9.f, 0.f, 15.75f}, 12.00f, 13.50f, 33.95f, 33.85f, {50.30f, 31.10f, 15.75f}, {-34.75f, 33.00f, 15.75f}},
        {{131.32f, 87.1704f, 31.8944f}, {9.00857f, 0.f, 12.0942f}, 12.50f, 17.00f, 31.9242f, 27.9242f, {49.97f, 27.80f, 12.0942f}, {-35.43f, 20.28f, 12.0942f}},
        {{133.992f, 83.021f, 32.8f}, {12.5f, 0.f, 11.75f}, 13.50f, 15.00f, 29.7f, 29.666f, {51.50f, 26.67f, 11.75f}, {-35.75f, 35.00f, 11.75f}},
        {{129.519f, 84.6879f, 36.6591f}, {13.8757f, 0.f, 20.755f}, 12.50f, 15.00f, 38.755f, 37.055f, {51.25f, 25.90f, 20.755f}, {-34.00f, 29.50f, 20.755f}},
        {{123.22f, 79.2103f, 44.1591f}, {11.3757f, 0.f, 21.505f}, 15.00f, 15.00f, 39.505f, 39.105f, {51.25f, 25.90f, 21.505f}, {-33.75f, 29.50f, 21.505f}},
    };
    return T[(i < 0 || i >= kNumCarPresets) ? 0 : i];
}
constexpr float DODGE_DEADZONE = 0.5f;  // CarConfig.h default
// Bullet
constexpr float BOX_MARGIN = 0.04f;          // btCollisionMargin.h:22 CONVEX_DISTANCE_MARGIN
constexpr float CONTACT_BREAKING = 0.02f;    // btPersistentManifold.cpp:25 gContactBreakingThreshold
constexpr float ERP2 = 0.8f;                 // Arena.cpp:486-488
constexpr float RESTITUTION_VEL_THRESH = 0.2f;  // btContactSolverInfo.h
}  // namespace C

// ---- controls ------------------------------------------------------------------------
struct Controls {
    float throttle, steer, pitch, yaw, roll;
    int32_t jump, boost, handbrake;
};

// ---- per-car state -------------------------------------------------------------------
struct CarS {
    // rigid body (Bullet units)
    V3 pos, vel, angvel;
    M3 rot;  // basis; columns forward/right/up
    // CarState
    int32_t isOnGround, wheelContact[4];
    int32_t hasJumped, hasDoubleJumped, hasFlipped, isFlipping, isJumping;
    V3 flipRelTorque;
    float jumpTime, flipTime, airTime, airTimeSinceJump;
    float boost, timeSpentBoosting;
    int32_t isSupersonic;
    float supersonicTime, handbrakeVal;
    int32_t isAutoFlipping;
    float autoFlipTimer, autoFlipTorqueScale;
    int32_t worldContactHas;
    V3 worldContactNormal;
    int32_t carContactOtherId;
    float carContactCooldown;
    int32_t isDemoed;
    float demoRespawnTimer;
    int32_t hitValid;
    V3 hitRelPos, hitBallPos, hitExtraVel;
    int32_t hitTickLo, hitTickHi, hitExtraTickLo, hitExtraTickHi;
    Controls lastControls;
    Controls controls;
    // btVehicleRL values that survive a tick
    float wheelSteer, wheelEngine, wheelBrake;
    float wheelLat[4], wheelLong[4], wheelPush[4];
    // gym layer: PlayerData match counters (Gym.cpp:6-38) + EventReward memo (CommonRewards.cpp:26-43)
    int32_t matchGoals, matchSaves, matchAssists, matchShots, matchShotPasses, matchBumps, matchDemos, boostPickups;
    float eventMemo[11];
    int32_t touchedStep;  // PlayerData::ballTouchedStep of the last snapshot
    int32_t snapIsDemoed; // PlayerData::carState.isDemoed of the last snapshot (read by Match::ParseActions)
    float prevAction[8];  // Match::prevActions row
};

struct BallS {
    V3 pos, vel, angvel;
    int32_t updateCounterLo;  // BallState::updateCounter (only compared, 32 bits suffice per episode)
};

// All 34 boost pads of an arena (R/Sim/BoostPad/BoostPad.h BoostPadState x 34) as bit masks + sparse cooldowns: a tick
// only touches the pads that are cooling down or that a car is near.
//   active  bit i : BoostPadState::isActive
//   cooling bit i : cooldown[i] > 0
//   locked byte i : prevLockedCarID (car index + 1, 0 = none)
struct PadsS {
    uint32_t activeLo, activeHi;
    uint32_t coolingLo, coolingHi;
    float cooldown[kNumPads];
    uint32_t locked[(kNumPads + 3) / 4];
};
constexpr uint64_t kAllPadsMask = (1ULL << kNumPads) - 1;
RL_HDI uint64_t pads_active(const PadsS& p) { return ((uint64_t)p.activeHi << 32) | p.activeLo; }
RL_HDI uint64_t pads_cooling(const PadsS& p) { return ((uint64_t)p.coolingHi << 32) | p.coolingLo; }
RL_HDI void pads_set_active(PadsS& p, uint64_t m) { p.activeLo = (uint32_t)m; p.activeHi = (uint32_t)(m >> 32); }
RL_HDI void pads_set_cooling(PadsS& p, uint64_t m) { p.coolingLo = (uint32_t)m; p.coolingHi = (uint32_t)(m >> 32); }
RL_HDI int pad_locked(const PadsS& p, int i) { return (int)((p.locked[i >> 2] >> ((i & 3) * 8)) & 0xffu); }
RL_HDI void pad_set_locked(PadsS& p, int i, int v) {
    uint32_t sh = (uint32_t)(i & 3) * 8;
    p.locked[i >> 2] = (p.locked[i >> 2] & ~(0xffu << sh)) | ((uint32_t)(v & 0xff) << sh);
}
RL_HDI void pads_reset(PadsS& p) {  // all active, no cooldown, nobody locked (Match.cpp:66-67 / fresh arena)
    pads_set_active(p, kAllPadsMask); pads_set_cooling(p, 0);
    for (int i = 0; i < kNumPads; i++) p.cooldown[i] = 0.f;
    for (int i = 0; i < (kNumPads + 3) / 4; i++) p.locked[i] = 0u;
}
RL_HDI void pad_set(PadsS& p, int i, bool isActive, float cooldown, int prevLockedCarId) {
    uint64_t bit = 1ULL << i;
    pads_set_active(p, isActive ? (pads_active(p) | bit) : (pads_active(p) & ~bit));
    pads_set_cooling(p, cooldown > 0 ? (pads_cooling(p) | bit) : (pads_cooling(p) & ~bit));
    p.cooldown[i] = cooldown;
    pad_set_locked(p, i, prevLockedCarId);
}
RL_HDI int lowest_bit(uint64_t m) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)m) - 1;
#else
    return __builtin_ctzll(m);
#endif
}

struct ArenaS {
    BallS ball;
    PadsS pads;
    int32_t tickLo, tickHi;  // Arena::tickCount
    // GameEventTracker persistent info
    float shotCooldown;
    int32_t ballShot, ballShotGoalTeam, ballScoredLast, lastBallUpdateCount;
    // GameState
    int32_t scoreLine[2], lastTouchCarId;
    int32_t lastTickLo, lastTickHi;  // GameState::lastTickCount
    // terminal conditions
    int32_t stepsSinceTouch;
    // RNG (pcg32 state), keyed by global arena id
    uint32_t rngLo, rngHi;
    CarS cars[kMaxCars];
};

constexpr int kCarWords = sizeof(CarS) / 4;
constexpr int kArenaHeaderWords = (sizeof(ArenaS) - sizeof(CarS) * kMaxCars) / 4;
static_assert(sizeof(CarS) % 4 == 0 && sizeof(ArenaS) % 4 == 0, "state must be word addressable");

RL_HDI int arena_words(int ncars) { return kArenaHeaderWords + ncars * kCarWords; }

RL_HDI int64_t get_i64(int32_t lo, int32_t hi) { return (int64_t)(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo); }
RL_HDI void set_i64(int32_t& lo, int32_t& hi, int64_t v) { lo = (int32_t)(uint32_t)((uint64_t)v & 0xffffffffu); hi = (int32_t)(uint32_t)((uint64_t)v >> 32); }

// team of car index c (id = c+1): Gym::Gym adds BLUE then ORANGE per team slot (Gym.cpp:46-50)
RL_HDI int car_team(int c, int spawnOpponents) { return spawnOpponents ? (c & 1) : 0; }

// pcg32
RL_HD RL_NOINLINE inline uint32_t rng_next(ArenaS& a) {
    uint64_t s = ((uint64_t)a.rngHi << 32) | a.rngLo;
    uint64_t old = s;
    s = old * 6364136223846793005ULL + 1442695040888963407ULL;
    a.rngLo = (uint32_t)s; a.rngHi = (uint32_t)(s >> 32);
    uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
}
// Math::RandFloat(min,max) = min + (r / float(max_r)) * (max - min)  (reference Math.cpp:54-57)
RL_HDI float rng_float(ArenaS& a, float lo, float hi) {
    float u = s_mul((float)(rng_next(a) >> 8), 1.0f / 16777215.0f);
    return s_add(lo, s_mul(u, s_sub(hi, lo)));
}

// static configuration shared by all arenas of an engine (kernel parameter, by value)
struct RewardTerm { int32_t kind; float weight; float params[11]; };
struct Mut {  // MutatorConfig as the tick reads it (rlg_mutators; Bullet units where the use site wants them)
    V3 gravityBT;               // world gravity, uu/s^2 * UU2BT
    float gravityX, gravityZ;   // uu/s^2 (Arena::IsBallProbablyGoingIn)
    float carWorldFriction, carWorldRestitution, ballWorldFriction, ballWorldRestitution;
    float ballMaxSpeed, jumpAccel, jumpImmediateForce, boostAccelGround, boostAccelAir, boostUsedPerSecond;
    float respawnDelay, bumpCooldownTime, padCooldownBig, padCooldownSmall, carSpawnBoost, ballHitExtraForceScale, bumpForceScale;
    float goalBaseThresholdY;
    float ballMass, ballRadius;  // Bullet mass units, uu (the ball's rigid body and sphere are built from them: Ball.cpp:74-91)
    int32_t unlimitedFlips, unlimitedDoubleJumps, demoMode, enableTeamDemos;
};
struct SimCfg {
    Mut mut;
    int32_t numArenas, numCars, spawnOpponents, tickSkip;
    int32_t numActions;  // rows of Tables::actions (90 unless rlg_engine_set_action_table replaced the table)
    int32_t carPreset;  // rlg_engine_cfg.car_preset
    int32_t obsKind, obsMaxPlayers, obsSize;
    int32_t numRewardTerms;
    RewardTerm rewards[8];
    int32_t zeroSum; float teamSpirit, opponentScale;
    int32_t noTouchMaxSteps, goalScoreTerminal;
    int32_t stateSetter, randBallSpeed, randCarSpeed, carsOnGround;
    int32_t playerOrder[kMaxCars];  // players[i] -> car index
    float ballDampFactor;           // powf(1 - BALL_DRAG, dt) computed on the host like btRigidBody::applyDamping
    float flipZDampFactor;          // powf(1 - FLIP_Z_DAMP_120, 1) (Car.cpp:753)
};

}  // namespace rl
