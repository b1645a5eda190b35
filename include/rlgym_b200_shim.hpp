// rlgym_b200_shim.hpp — header-only C++ host layer over the C ABI (rlgym_b200.h) that keeps the reference's
// plugin / operator surface for the collection path, so user code written against RLGymSim_CPP / RLGymPPO_CPP
// (T/examplemain.cpp:58-100) keeps compiling:
//
//   RLGSC::  Vec, RotMat, PhysObj, Action, FList/FList2/IList, CarState, PlayerData, GameState,
//            OBSBuilder, RewardFunction, ActionParser, StateSetter, TerminalCondition (virtual bases, same virtuals:
//            G/Utils/OBSBuilders/OBSBuilder.h:10-15, RewardFunctions/RewardFunction.h:9-35, ActionParsers/ActionParser.h:11-14,
//            StateSetters/StateSetter.h:9, TerminalConditions/TerminalCondition.h:7-8),
//            built-ins DefaultOBS, DefaultOBSPadded, CombinedReward, EventReward(WeightScales), VelocityPlayerToBallReward,
//            VelocityBallToGoalReward, FaceBallReward, VelocityReward, ZeroSumReward, DiscreteAction, NoTouchCondition,
//            GoalScoreCondition, KickoffState, RandomState, Match (G/Envs/Match.h:27-46), Gym (G/Gym.h:8-31)
//   RLGPC::  PPOLearnerConfig, LearnerConfig (field for field), Report, AvgTracker, Timer, WelfordRunningStat, StepCallback,
//            EnvCreateResult / EnvCreateFn, GameInst (GameInst.h:7-60), GameTrajectory (device tensors in the reference's row
//            order), DiscretePolicy / ExperienceBuffer handles, ThreadAgentManager (ThreadAgentManager.h:10-60: same constructor,
//            CreateAgents / StartAgents / StopAgents / SetStepCallback / CollectTimesteps -> GameTrajectory / GetMetrics /
//            ResetMetrics), PPOLearner (csrc/ppo.cu), Learner(EnvCreateFn, LearnerConfig) + Learn() (Learner.h:14-60): the
//            reference's examplemain.cpp compiles against this header unmodified (tests/test_cpp_shim.py)
//   RocketSim::Init(path)  — loads soccar/*.cmf exactly like R/RocketSim.cpp:70-212 and keeps the bytes for the engines
//
// Built-in plugins are recognised by dynamic_cast and become configuration of the fused device kernels
// (RLGB200::PlanFromMatch); with only built-ins and no StepCallback nothing runs on the host.  Anything user-defined takes the
// HOST-PLUGIN path for that stage only: every env-step the snapshot GameState of every arena (taken where G/Gym.cpp:84-93
// takes it) comes to the host in one batched copy while the GPU runs the step's remaining ticks, GameState / PlayerData are
// filled (incl. inverted copies, pads in CommonValues order, match counters) and the user's virtuals are called from
// LearnerConfig::numThreads host threads — one Match per arena from EnvCreateFn like the reference — and the rows they produce
// are uploaded into the trajectory ring:
//   user OBSBuilder          -> obs rows built on the host (Reset / PreStep / BuildOBS)
//   user RewardFunction      -> the whole reward graph on the host (built-in reward classes have host implementations too, so a
//                               CombinedReward may mix stock and user terms; ZeroSumReward likewise)
//   user TerminalCondition   -> OR-ed with the fused built-in conditions
//   StepCallback             -> called per arena per step with a filled Gym::StepResult (GameInst.cpp:23-24)
//   user StateSetter         -> Arena / Car / Ball proxies -> rlg_engine_set_state + rlg_engine_reset_current
//   user ActionParser        -> probed once into its table (index -> Action for every index below GetActionAmount() <= 96) and uploaded
//                               with rlg_engine_set_action_table; the policy head gets GetActionAmount() outputs.  The mapping must not depend
//                               on the game state (checked on two different probe states; a state-dependent parser is refused loudly).
// Errors: every failing C-ABI call is re-thrown as std::runtime_error(rlg_last_error()) like RG_ERR_CLOSE
// (G/Framework.h:17-22).
#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <ctime>
#include <dirent.h>
#include <exception>
#include <fstream>
#include <functional>
#include <iomanip>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <typeinfo>
#include <unistd.h>
#include <utility>
#include <vector>

#include "rlgym_b200.h"

namespace RLGB200 {
inline void Check(int rc) {
    if (rc != RLG_OK) throw std::runtime_error(std::string("RLGB200: ") + rlg_last_error());
}
}  // namespace RLGB200

// ---------------------------------------------------------------------------------------------------------------------
namespace RocketSim {
// R/Sim/Car/CarConfig/CarConfig.h: the six stock configurations; the engine holds their geometry (rlg_engine_cfg.car_preset)
struct CarConfig { int preset; };
static const CarConfig CAR_CONFIG_OCTANE{RLG_CAR_OCTANE}, CAR_CONFIG_DOMINUS{RLG_CAR_DOMINUS}, CAR_CONFIG_PLANK{RLG_CAR_PLANK},
    CAR_CONFIG_BREAKOUT{RLG_CAR_BREAKOUT}, CAR_CONFIG_HYBRID{RLG_CAR_HYBRID}, CAR_CONFIG_MERC{RLG_CAR_MERC};
struct Vec {
    float x = 0, y = 0, z = 0, _w = 0;  // R/Math/MathTypes/MathTypes.h:7-16 (16-byte Vec)
    Vec() = default;
    Vec(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    Vec operator+(const Vec& o) const { return Vec(x + o.x, y + o.y, z + o.z); }
    Vec operator-(const Vec& o) const { return Vec(x - o.x, y - o.y, z - o.z); }
    Vec operator*(float s) const { return Vec(x * s, y * s, z * s); }
    Vec operator*(const Vec& o) const { return Vec(x * o.x, y * o.y, z * o.z); }
    Vec operator/(float s) const { return Vec(x / s, y / s, z / s); }
    Vec operator-() const { return Vec(-x, -y, -z); }
    Vec& operator+=(const Vec& o) { x += o.x; y += o.y; z += o.z; return *this; }
    Vec& operator-=(const Vec& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    Vec& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
    float& operator[](size_t i) { return (&x)[i]; }
    float operator[](size_t i) const { return (&x)[i]; }
    float Dot(const Vec& o) const { return x * o.x + y * o.y + z * o.z + _w * o._w; }
    Vec Cross(const Vec& o) const { return Vec(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
    float LengthSq() const { return Dot(*this); }
    float LengthSq2D() const { return x * x + y * y; }
    float Length() const { float l = LengthSq(); return l > 0 ? std::sqrt(l) : 0; }  // R/Math/MathTypes/MathTypes.h: sqrtf when > 0
    float Length2D() const { float l = LengthSq2D(); return l > 0 ? std::sqrt(l) : 0; }
    float Dist(const Vec& o) const { return (*this - o).Length(); }
    float Dist2D(const Vec& o) const { return (*this - o).Length2D(); }
    // MathTypes.h Vec::Normalized: (*this) / Length() when the squared length exceeds FLT_EPSILON^2, else the zero vector
    Vec Normalized() const { float l = Length(); return l > 1.1920929e-07f * 1.1920929e-07f ? Vec(x / l, y / l, z / l) : Vec(); }
};
// R/Sim/GameMode.h, R/Sim/MutatorConfig/MutatorConfig.h:10-72: same field names and defaults; Gym's last constructor arguments.
// Only soccar is built (north_star names soccar); carMass / ballMass / ballRadius must keep their defaults.
enum class GameMode : uint8_t { SOCCAR = 0 };
enum class DemoMode : uint8_t { NORMAL = RLG_DEMO_NORMAL, ON_CONTACT = RLG_DEMO_ON_CONTACT, DISABLED = RLG_DEMO_DISABLED };
struct MutatorConfig {
    Vec gravity;
    float carMass, carWorldFriction, carWorldRestitution, ballMass, ballMaxSpeed, ballDrag, ballWorldFriction, ballWorldRestitution, jumpAccel,
        jumpImmediateForce, boostAccelGround, boostAccelAir, boostUsedPerSecond, respawnDelay, bumpCooldownTime, boostPadCooldown_Big,
        boostPadCooldown_Small, carSpawnBoostAmount;
    float ballHitExtraForceScale = 1, bumpForceScale = 1;
    float ballRadius;
    bool unlimitedFlips = false, unlimitedDoubleJumps = false;
    DemoMode demoMode = DemoMode::NORMAL;
    bool enableTeamDemos = false;
    float goalBaseThresholdY;
    MutatorConfig(GameMode = GameMode::SOCCAR) {
        rlg_mutators m;
        rlg_mutators_default(&m);
        gravity = Vec(m.gravity[0], m.gravity[1], m.gravity[2]);
        carMass = m.car_mass; carWorldFriction = m.car_world_friction; carWorldRestitution = m.car_world_restitution;
        ballMass = m.ball_mass; ballMaxSpeed = m.ball_max_speed; ballDrag = m.ball_drag;
        ballWorldFriction = m.ball_world_friction; ballWorldRestitution = m.ball_world_restitution;
        jumpAccel = m.jump_accel; jumpImmediateForce = m.jump_immediate_force;
        boostAccelGround = m.boost_accel_ground; boostAccelAir = m.boost_accel_air; boostUsedPerSecond = m.boost_used_per_second;
        respawnDelay = m.respawn_delay; bumpCooldownTime = m.bump_cooldown_time;
        boostPadCooldown_Big = m.boost_pad_cooldown_big; boostPadCooldown_Small = m.boost_pad_cooldown_small;
        carSpawnBoostAmount = m.car_spawn_boost_amount; ballRadius = m.ball_radius; goalBaseThresholdY = m.goal_base_threshold_y;
    }
    rlg_mutators ToC() const {  // field for field
        rlg_mutators m;
        m.gravity[0] = gravity.x; m.gravity[1] = gravity.y; m.gravity[2] = gravity.z;
        m.car_mass = carMass; m.car_world_friction = carWorldFriction; m.car_world_restitution = carWorldRestitution;
        m.ball_mass = ballMass; m.ball_max_speed = ballMaxSpeed; m.ball_drag = ballDrag;
        m.ball_world_friction = ballWorldFriction; m.ball_world_restitution = ballWorldRestitution;
        m.jump_accel = jumpAccel; m.jump_immediate_force = jumpImmediateForce;
        m.boost_accel_ground = boostAccelGround; m.boost_accel_air = boostAccelAir; m.boost_used_per_second = boostUsedPerSecond;
        m.respawn_delay = respawnDelay; m.bump_cooldown_time = bumpCooldownTime;
        m.boost_pad_cooldown_big = boostPadCooldown_Big; m.boost_pad_cooldown_small = boostPadCooldown_Small;
        m.car_spawn_boost_amount = carSpawnBoostAmount; m.ball_hit_extra_force_scale = ballHitExtraForceScale; m.bump_force_scale = bumpForceScale;
        m.ball_radius = ballRadius; m.unlimited_flips = unlimitedFlips; m.unlimited_double_jumps = unlimitedDoubleJumps;
        m.demo_mode = (int32_t)demoMode; m.enable_team_demos = enableTeamDemos; m.goal_base_threshold_y = goalBaseThresholdY;
        return m;
    }
};
struct RotMat {
    Vec forward{1, 0, 0}, right{0, 1, 0}, up{0, 0, 1};
};
struct Angle {
    float yaw = 0, pitch = 0, roll = 0;
    Angle() = default;
    Angle(float y, float p, float r) : yaw(y), pitch(p), roll(r) {}
    // R/Math/MathTypes/MathTypes.cpp:73-78 -> btMatrix3x3::setEulerYPR(yaw, -pitch, -roll)
    RotMat ToRotMat() const {
        float eulerX = -roll, eulerY = -pitch, eulerZ = yaw;
        float ci = std::cos(eulerX), cj = std::cos(eulerY), ch = std::cos(eulerZ);
        float si = std::sin(eulerX), sj = std::sin(eulerY), sh = std::sin(eulerZ);
        float cc = ci * ch, cs = ci * sh, sc = si * ch, ss = si * sh;
        RotMat m;  // columns of the Bullet basis
        m.forward = Vec(cj * ch, cj * sh, -sj);
        m.right = Vec(sj * sc - cs, sj * ss + cc, cj * si);
        m.up = Vec(sj * cc + ss, sj * cs - sc, cj * ci);
        return m;
    }
};
enum class Team : uint8_t { BLUE = 0, ORANGE = 1 };

struct CarControls {
    float throttle = 0, steer = 0, pitch = 0, yaw = 0, roll = 0;
    bool jump = false, boost = false, handbrake = false;
};

struct BallHitInfo {  // R/Sim/BallHitInfo/BallHitInfo.h
    bool isValid = false;
    Vec relativePosOnBall, ballPos, extraHitVel;
    uint64_t tickCountWhenHit = ~0ULL, tickCountWhenExtraImpulseApplied = ~0ULL;
};
// R/Sim/Car/Car.h:17-115, same member names and defaults
struct CarState {
    Vec pos{0, 0, 17.f};
    RotMat rotMat;
    Vec vel, angVel;
    bool isOnGround = true;
    bool wheelsWithContact[4] = {};
    bool hasJumped = false, hasDoubleJumped = false, hasFlipped = false;
    Vec flipRelTorque;
    float jumpTime = 0, flipTime = 0;
    bool isFlipping = false, isJumping = false;
    float airTime = 0, airTimeSinceJump = 0;
    float boost = 100.f / 3, timeSpentBoosting = 0;
    bool isSupersonic = false;
    float supersonicTime = 0, handbrakeVal = 0;
    bool isAutoFlipping = false;
    float autoFlipTimer = 0, autoFlipTorqueScale = 0;
    struct { bool hasContact = false; Vec contactNormal; } worldContact;
    struct { uint32_t otherCarID = 0; float cooldownTimer = 0; } carContact;
    bool isDemoed = false;
    float demoRespawnTimer = 0;
    BallHitInfo ballHitInfo;
    CarControls lastControls;
};
struct BallState {
    Vec pos{0, 0, 93.15f}, vel, angVel;
};

// Proxies handed to StateSetter::ResetState(Arena*): SetState stages into host arrays that the engine uploads.
class Car {
public:
    uint32_t id = 0;
    Team team = Team::BLUE;
    CarState GetState() const { return state; }
    void SetState(const CarState& s) { state = s; dirty = true; }
    CarState state;
    bool dirty = false;
};
class Ball {
public:
    BallState GetState() const { return state; }
    void SetState(const BallState& s) { state = s; dirty = true; }
    BallState state;
    bool dirty = false;
};
class Arena {
public:
    std::vector<Car*> _cars;
    Ball* ball = nullptr;
    const std::vector<Car*>& GetCars() const { return _cars; }
    // Arena::ResetToRandomKickoff (R/Sim/Arena/Arena.cpp:112-216) on the device for this arena
    void ResetToRandomKickoff(int = -1) { wantsKickoff = true; }
    bool wantsKickoff = false;
};

// RocketSim::Init (R/RocketSim.cpp:70-212): reads <path>/soccar/*.cmf; the blobs are handed to every engine.
inline std::vector<std::string>& CollisionMeshBlobs() {
    static std::vector<std::string> blobs;
    return blobs;
}
inline void Init(const std::string& collisionMeshesFolder) {
    auto& blobs = CollisionMeshBlobs();
    blobs.clear();
    std::string dir = collisionMeshesFolder + "/soccar";
    std::vector<std::string> names;
    if (DIR* d = opendir(dir.c_str())) {
        while (dirent* ent = readdir(d)) {
            std::string n = ent->d_name;
            if (n.size() > 4 && n.substr(n.size() - 4) == ".cmf") names.push_back(n);
        }
        closedir(d);
    }
    std::sort(names.begin(), names.end());
    for (auto& n : names) {
        std::ifstream f(dir + "/" + n, std::ios::binary);
        blobs.emplace_back((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    }
    if (blobs.empty()) throw std::runtime_error("RocketSim::Init: no collision meshes found in " + dir + " (R/Sim/Arena/Arena.cpp:1021-1026)");
}
}  // namespace RocketSim

// ---------------------------------------------------------------------------------------------------------------------
using namespace RocketSim;  // G/Framework.h:8 does the same: user code names Vec, CarConfig, MutatorConfig unqualified
namespace RLGSC {
using RocketSim::Vec; using RocketSim::RotMat; using RocketSim::Angle; using RocketSim::Team; using RocketSim::CarState;
using RocketSim::BallState; using RocketSim::Arena; using RocketSim::Car; using RocketSim::Ball;
typedef std::vector<float> FList;
typedef std::vector<FList> FList2;
typedef std::vector<int> IList;

// G/Lists.h: obs builders append with +=
inline FList& operator+=(FList& l, float v) { l.push_back(v); return l; }
inline FList& operator+=(FList& l, const Vec& v) { l.push_back(v.x); l.push_back(v.y); l.push_back(v.z); return l; }
inline FList& operator+=(FList& l, const FList& o) { l.insert(l.end(), o.begin(), o.end()); return l; }
inline FList& operator+=(FList& l, std::initializer_list<float> o) { l.insert(l.end(), o.begin(), o.end()); return l; }

namespace CommonValues {  // G/Utils/CommonValues.h
constexpr float SIDE_WALL_X = 4096, BACK_WALL_Y = 5120, CEILING_Z = 2044, BACK_NET_Y = 6000, GOAL_HEIGHT = 642.775f, GRAVITY_Z = -650.f,
                BOOST_CONSUMED_PER_SECOND = 100.f / 3.f;
constexpr float BALL_RADIUS = 92.75f, BALL_MAX_SPEED = 6000, CAR_MAX_SPEED = 2300, SUPERSONIC_THRESHOLD = 2200, CAR_MAX_ANG_VEL = 5.5f;
constexpr float BLUE_TEAM = 0, ORANGE_TEAM = 1, NUM_ACTIONS = 8;
constexpr int BOOST_LOCATIONS_AMOUNT = RLG_NUM_PADS;
inline Vec ORANGE_GOAL_CENTER() { return Vec(0, BACK_WALL_Y, GOAL_HEIGHT / 2); }
inline Vec BLUE_GOAL_CENTER() { return Vec(0, -BACK_WALL_Y, GOAL_HEIGHT / 2); }
static const Vec ORANGE_GOAL_BACK(0, BACK_NET_Y, GOAL_HEIGHT / 2), BLUE_GOAL_BACK(0, -BACK_NET_Y, GOAL_HEIGHT / 2);
}  // namespace CommonValues
namespace Math {
// G/Math.cpp:3-5: |y| > SOCCAR_GOAL_SCORE_BASE_THRESHOLD_Y + BALL_COLLISION_RADIUS_SOCCAR
inline bool IsBallScored(Vec pos) { return std::fabs(pos.y) > 5124.25f + 91.25f; }
}  // namespace Math

struct Action {  // G/Utils/BasicTypes/Action.h:5-47
    float throttle = 0, steer = 0, pitch = 0, yaw = 0, roll = 0, jump = 0, boost = 0, handbrake = 0;
    constexpr static int ELEM_AMOUNT = 8;
    float operator[](size_t i) const { return (&throttle)[i]; }
    float& operator[](size_t i) { return (&throttle)[i]; }
};
typedef std::vector<Action> ActionSet;

struct PhysObj {  // G/Utils/Gamestates/PhysObj.h
    Vec pos, vel, angVel;
    RotMat rotMat;
    PhysObj() = default;
    explicit PhysObj(const CarState& s) : pos(s.pos), vel(s.vel), angVel(s.angVel), rotMat(s.rotMat) {}
    explicit PhysObj(const BallState& s) : pos(s.pos), vel(s.vel), angVel(s.angVel) {}
    PhysObj Invert() const {  // PhysObj.cpp:19-31
        const Vec inv(-1, -1, 1);
        PhysObj r = *this;
        r.pos = pos * inv; r.vel = vel * inv; r.angVel = angVel * inv;
        r.rotMat.forward = rotMat.forward * inv; r.rotMat.right = rotMat.right * inv; r.rotMat.up = rotMat.up * inv;
        return r;
    }
};
struct PlayerData {  // G/Utils/Gamestates/PlayerData.h:7-38
    uint32_t carId = 0;
    Team team = Team::BLUE;
    PhysObj phys, physInv;
    CarState carState;
    int matchGoals = 0, matchSaves = 0, matchAssists = 0, matchShots = 0, matchShotPasses = 0, matchBumps = 0, matchDemos = 0, boostPickups = 0;
    bool hasJump = false, hasFlip = false;
    float boostFraction = 0;
    bool ballTouchedStep = false, ballTouchedTick = false;
    const PhysObj& GetPhys(bool inverted) const { return inverted ? physInv : phys; }
};
struct ScoreLine {
    int teamGoals[2] = {0, 0};
    int operator[](size_t i) const { return teamGoals[i]; }
    int& operator[](size_t i) { return teamGoals[i]; }
};
struct GameState {  // G/Utils/Gamestates/GameState.h:19-57
    float deltaTime = 0;
    ScoreLine scoreLine;
    int lastTouchCarID = -1;
    std::vector<PlayerData> players;
    BallState ballState;
    PhysObj ball, ballInv;
    std::array<bool, RLG_NUM_PADS> boostPads{}, boostPadsInv{};
    std::array<float, RLG_NUM_PADS> boostPadTimers{}, boostPadTimersInv{};
    Arena* lastArena = nullptr;  // "could be null" in the reference too; always null here (the arena lives on the device)
    uint64_t lastTickCount = 0;
    const PhysObj& GetBallPhys(bool inverted) const { return inverted ? ballInv : ball; }
    const std::array<bool, RLG_NUM_PADS>& GetBoostPads(bool inverted) const { return inverted ? boostPadsInv : boostPads; }
};

class OBSBuilder {
public:
    virtual void Reset(const GameState&) {}
    virtual void PreStep(const GameState&) {}
    virtual FList BuildOBS(const PlayerData& player, const GameState& state, const Action& prevAction) = 0;
    virtual ~OBSBuilder() = default;
};
class RewardFunction {
public:
    virtual void Reset(const GameState&) {}
    virtual void PreStep(const GameState&) {}
    virtual float GetReward(const PlayerData&, const GameState&, const Action&) { throw std::runtime_error("GetReward() is unimplemented"); }
    virtual float GetFinalReward(const PlayerData& p, const GameState& s, const Action& a) { return GetReward(p, s, a); }
    virtual std::vector<float> GetAllRewards(const GameState& state, const ActionSet& prevActions, bool final) {
        std::vector<float> r(state.players.size());
        for (size_t i = 0; i < r.size(); i++) r[i] = final ? GetFinalReward(state.players[i], state, prevActions[i]) : GetReward(state.players[i], state, prevActions[i]);
        return r;
    }
    virtual ~RewardFunction() = default;
};
class ActionParser {
public:
    typedef IList Input;
    virtual ActionSet ParseActions(const IList& actionsData, const GameState& gameState) = 0;
    virtual int GetActionAmount() = 0;
    virtual ~ActionParser() = default;
};
class StateSetter {
public:
    virtual GameState ResetState(Arena* arena) = 0;
    virtual ~StateSetter() = default;
};
class TerminalCondition {
public:
    virtual void Reset(const GameState&) {}
    virtual bool IsTerminal(const GameState& currentState) = 0;
    virtual ~TerminalCondition() = default;
};

// ---- built-ins ----------------------------------------------------------------------------------------------------
// Configuration carriers for the fused device kernels.  The reward and terminal classes ALSO keep working host virtuals
// (same arithmetic as G/Utils/RewardFunctions/CommonRewards.h / TerminalConditions/*.h) so that a user's CombinedReward may mix
// them with user-defined terms on the host-plugin path; the obs builders and state setters run on the device only.
#define RLGB200_DEVICE_ONLY(what) throw std::runtime_error(std::string(what) + ": built-in plugin runs fused on the device; its host virtual is not called")
class DefaultOBS : public OBSBuilder {  // G/Utils/OBSBuilders/DefaultOBS.h
public:
    FList BuildOBS(const PlayerData&, const GameState&, const Action&) override { RLGB200_DEVICE_ONLY("DefaultOBS"); }
};
class DefaultOBSPadded : public OBSBuilder {  // DefaultOBSPadded.h
public:
    int maxPlayers;
    explicit DefaultOBSPadded(int maxPlayers_) : maxPlayers(maxPlayers_) {}
    FList BuildOBS(const PlayerData&, const GameState&, const Action&) override { RLGB200_DEVICE_ONLY("DefaultOBSPadded"); }
};
class EventReward : public RewardFunction {  // CommonRewards.h:6-49, CommonRewards.cpp
public:
    struct WeightScales {
        float goal = 0, teamGoal = 0, concede = 0, assist = 0, touch = 0, shot = 0, shotPass = 0, save = 0, demo = 0, demoed = 0, boostPickup = 0;
        float& operator[](size_t i) { return (&goal)[i]; }
        float operator[](size_t i) const { return (&goal)[i]; }
    };
    WeightScales weights;
    std::map<uint32_t, std::array<float, 11>> lastRegisteredValues;
    explicit EventReward(WeightScales w) : weights(w) {}
    static std::array<float, 11> ExtractValues(const PlayerData& p, const GameState& s) {
        const int own = (int)p.team;
        return {(float)p.matchGoals, (float)s.scoreLine[own], (float)s.scoreLine[1 - own], (float)p.matchAssists, (float)p.ballTouchedStep, (float)p.matchShots,
                (float)p.matchShotPasses, (float)p.matchSaves, (float)p.matchDemos, (float)p.carState.isDemoed, p.boostFraction};
    }
    void Reset(const GameState& s) override {
        lastRegisteredValues.clear();
        for (auto& p : s.players) lastRegisteredValues[p.carId] = ExtractValues(p, s);
    }
    float GetReward(const PlayerData& p, const GameState& s, const Action&) override {
        auto& old = lastRegisteredValues[p.carId];
        const auto cur = ExtractValues(p, s);
        float r = 0;
        for (int i = 0; i < 11; i++) r += std::max(cur[i] - old[i], 0.f) * weights[i];
        old = cur;
        return r;
    }
};
class VelocityReward : public RewardFunction {
public:
    bool isNegative;
    explicit VelocityReward(bool isNegative_ = false) : isNegative(isNegative_) {}
    float GetReward(const PlayerData& p, const GameState&, const Action&) override { return p.phys.vel.Length() / CommonValues::CAR_MAX_SPEED * (1 - 2 * isNegative); }
};
class SaveBoostReward : public RewardFunction {  // CommonRewards.h:61-70
public:
    float exponent;
    explicit SaveBoostReward(float exponent_ = 0.5f) : exponent(exponent_) {}
    float GetReward(const PlayerData& p, const GameState&, const Action&) override { return std::min(std::max(powf(p.boostFraction, exponent), 0.f), 1.f); }
};
class TouchBallReward : public RewardFunction {  // CommonRewards.h:110-124
public:
    float aerialWeight;
    explicit TouchBallReward(float aerialWeight_ = 0) : aerialWeight(aerialWeight_) {}
    float GetReward(const PlayerData& p, const GameState& s, const Action&) override {
        return p.ballTouchedStep ? powf((s.ball.pos.z + CommonValues::BALL_RADIUS) / (CommonValues::BALL_RADIUS * 2), aerialWeight) : 0.f;
    }
};
class VelocityPlayerToBallReward : public RewardFunction {
public:
    float GetReward(const PlayerData& p, const GameState& s, const Action&) override {
        return (s.ball.pos - p.phys.pos).Normalized().Dot(p.phys.vel / CommonValues::CAR_MAX_SPEED);
    }
};
class FaceBallReward : public RewardFunction {
public:
    float GetReward(const PlayerData& p, const GameState& s, const Action&) override { return p.carState.rotMat.forward.Dot((s.ball.pos - p.phys.pos).Normalized()); }
};
class VelocityBallToGoalReward : public RewardFunction {
public:
    bool ownGoal;
    explicit VelocityBallToGoalReward(bool ownGoal_ = false) : ownGoal(ownGoal_) {}
    float GetReward(const PlayerData& p, const GameState& s, const Action&) override {
        const bool orange = (p.team == Team::BLUE) != ownGoal;
        const Vec target = orange ? CommonValues::ORANGE_GOAL_BACK : CommonValues::BLUE_GOAL_BACK;
        return (target - s.ball.pos).Normalized().Dot(s.ball.vel / CommonValues::BALL_MAX_SPEED);
    }
};
class CombinedReward : public RewardFunction {  // CombinedReward.h
public:
    std::vector<RewardFunction*> rewardFuncs;
    std::vector<float> rewardWeights;
    bool ownsFuncs;
    CombinedReward(std::vector<RewardFunction*> funcs, std::vector<float> weights, bool ownsFuncs_ = false)
        : rewardFuncs(std::move(funcs)), rewardWeights(std::move(weights)), ownsFuncs(ownsFuncs_) {}
    CombinedReward(std::vector<std::pair<RewardFunction*, float>> funcsWithWeights, bool ownsFuncs_ = false) : ownsFuncs(ownsFuncs_) {
        for (auto& p : funcsWithWeights) { rewardFuncs.push_back(p.first); rewardWeights.push_back(p.second); }
    }
    void Reset(const GameState& s) override { for (auto* f : rewardFuncs) f->Reset(s); }
    void PreStep(const GameState& s) override { for (auto* f : rewardFuncs) f->PreStep(s); }
    std::vector<float> GetAllRewards(const GameState& state, const ActionSet& prev, bool final) override {
        std::vector<float> all(state.players.size());
        for (size_t i = 0; i < rewardFuncs.size(); i++) {
            const auto r = rewardFuncs[i]->GetAllRewards(state, prev, final);
            for (size_t j = 0; j < r.size(); j++) all[j] += r[j] * rewardWeights[i];
        }
        return all;
    }
    ~CombinedReward() override { if (ownsFuncs) for (auto* f : rewardFuncs) delete f; }
};
class ZeroSumReward : public RewardFunction {  // ZeroSumReward.h:8-26, ZeroSumReward.cpp:3-29
public:
    RewardFunction* childFunc;
    bool ownsFunc;
    float teamSpirit, opponentScale;
    ZeroSumReward(RewardFunction* child, float teamSpirit_, float opponentScale_ = 1, bool ownsFunc_ = false)
        : childFunc(child), ownsFunc(ownsFunc_), teamSpirit(teamSpirit_), opponentScale(opponentScale_) {}
    void Reset(const GameState& s) override { childFunc->Reset(s); }
    void PreStep(const GameState& s) override { childFunc->PreStep(s); }
    std::vector<float> GetAllRewards(const GameState& state, const ActionSet& prev, bool final) override {
        std::vector<float> r = childFunc->GetAllRewards(state, prev, final);
        int count[2] = {0, 0};
        float avg[2] = {0, 0};
        for (size_t i = 0; i < r.size(); i++) { const int t = (int)state.players[i].team; count[t]++; avg[t] += r[i]; }
        for (int t = 0; t < 2; t++) avg[t] /= (float)std::max(count[t], 1);
        for (size_t i = 0; i < r.size(); i++) {
            const int t = (int)state.players[i].team;
            r[i] = r[i] * (1 - teamSpirit) + (avg[t] * teamSpirit) - (avg[1 - t] * opponentScale);
        }
        return r;
    }
    ~ZeroSumReward() override { if (ownsFunc) delete childFunc; }
};
class DiscreteAction : public ActionParser {  // DiscreteAction.h:14
public:
    ActionSet ParseActions(const IList& idx, const GameState&) override {
        float t[RLG_NUM_ACTIONS * 8];
        RLGB200::Check(rlg_action_table(t));
        ActionSet out;
        for (int i : idx) { const float* r = t + 8 * i; out.push_back(Action{r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]}); }
        return out;
    }
    int GetActionAmount() override { return RLG_NUM_ACTIONS; }
};
class NoTouchCondition : public TerminalCondition {
public:
    int stepsSinceTouch = 0;
    int64_t maxSteps;
    explicit NoTouchCondition(int64_t maxSteps_) : maxSteps(maxSteps_) {}
    void Reset(const GameState&) override { stepsSinceTouch = 0; }
    bool IsTerminal(const GameState& s) override {
        for (auto& p : s.players) if (p.ballTouchedStep) { stepsSinceTouch = 0; return false; }
        return ++stepsSinceTouch >= maxSteps;
    }
};
class GoalScoreCondition : public TerminalCondition {
public:
    bool IsTerminal(const GameState& s) override { return Math::IsBallScored(s.ball.pos); }
};
class KickoffState : public StateSetter {
public:
    GameState ResetState(Arena* a) override { a->ResetToRandomKickoff(); return GameState(); }
};
class RandomState : public StateSetter {  // RandomState.h
public:
    bool randBallSpeed, randCarSpeed, carsOnGround;
    RandomState(bool randBallSpeed_, bool randCarSpeed_, bool carsOnGround_) : randBallSpeed(randBallSpeed_), randCarSpeed(randCarSpeed_), carsOnGround(carsOnGround_) {}
    GameState ResetState(Arena*) override { RLGB200_DEVICE_ONLY("RandomState"); }
};

class Match {  // G/Envs/Match.h:27-46
public:
    RewardFunction* rewardFn;
    std::vector<TerminalCondition*> terminalConditions;
    OBSBuilder* obsBuilder;
    ActionParser* actionParser;
    StateSetter* stateSetter;
    int teamSize;
    bool spawnOpponents;
    int playerAmount;
    ActionSet prevActions;
    Match(RewardFunction* rewardFn_, std::vector<TerminalCondition*> terminalConditions_, OBSBuilder* obsBuilder_, ActionParser* actionParser_,
          StateSetter* stateSetter_, int teamSize_ = 1, bool spawnOpponents_ = true)
        : rewardFn(rewardFn_), terminalConditions(std::move(terminalConditions_)), obsBuilder(obsBuilder_), actionParser(actionParser_),
          stateSetter(stateSetter_), teamSize(teamSize_), spawnOpponents(spawnOpponents_), playerAmount(teamSize_ * (spawnOpponents_ ? 2 : 1)) {}
};
class Gym {  // G/Gym.h:8-31 — with the device engine a Gym is a description; stepping happens batched (ThreadAgentManager)
public:
    Match* match;
    int tickSkip;
    GameState prevState;
    int totalTicks = 0, totalSteps = 0;
    struct StepResult {
        FList2 obs;
        FList reward;
        bool done;
        GameState state;
    };
    RocketSim::CarConfig carConfig;
    RocketSim::MutatorConfig mutatorConfig;
    Gym(Match* match_, int tickSkip_, RocketSim::CarConfig carConfig_ = RocketSim::CAR_CONFIG_OCTANE,
        RocketSim::GameMode gameMode = RocketSim::GameMode::SOCCAR, RocketSim::MutatorConfig mutatorConfig_ = RocketSim::MutatorConfig(RocketSim::GameMode::SOCCAR))
        : match(match_), tickSkip(tickSkip_), carConfig(carConfig_), mutatorConfig(mutatorConfig_) {}  // G/Gym.h:18 (soccar only)
    virtual ~Gym() = default;
};
}  // namespace RLGSC

// ---------------------------------------------------------------------------------------------------------------------
namespace RLGB200 {
using namespace RLGSC;

// Which stages of a Match are fused on the device and which run through the host-plugin path.
struct PluginPlan {
    rlg_engine_cfg cfg;
    bool obsOnHost = false, rewardOnHost = false;
    int hostTerminals = 0;  // user TerminalConditions (the built-in ones stay fused and are OR-ed in)
    std::vector<float> actionTable;  // a user ActionParser's table, 8 floats per action index (empty: DiscreteAction)
    bool AnyHost() const { return obsOnHost || rewardOnHost || hostTerminals > 0; }
};

inline bool BuiltinRewardTerm(RewardFunction* f, float weight, rlg_reward_term& t) {
    memset(&t, 0, sizeof(t));
    t.weight = weight;
    if (auto* e = dynamic_cast<EventReward*>(f)) {
        t.kind = RLG_REW_EVENT;
        for (int k = 0; k < 11; k++) t.params[k] = e->weights[k];
    } else if (dynamic_cast<VelocityPlayerToBallReward*>(f)) t.kind = RLG_REW_VEL_PLAYER_TO_BALL;
    else if (auto* g = dynamic_cast<VelocityBallToGoalReward*>(f)) { t.kind = RLG_REW_VEL_BALL_TO_GOAL; t.params[0] = g->ownGoal ? 1.f : 0.f; }
    else if (dynamic_cast<FaceBallReward*>(f)) t.kind = RLG_REW_FACE_BALL;
    else if (auto* v = dynamic_cast<VelocityReward*>(f)) { t.kind = RLG_REW_VELOCITY; t.params[0] = v->isNegative ? 1.f : 0.f; }
    else if (auto* sb = dynamic_cast<SaveBoostReward*>(f)) { t.kind = RLG_REW_SAVE_BOOST; t.params[0] = sb->exponent; }
    else if (auto* tb = dynamic_cast<TouchBallReward*>(f)) { t.kind = RLG_REW_TOUCH_BALL; t.params[0] = tb->aerialWeight; }
    else return false;
    // a SUBCLASS of a built-in (e.g. an EventReward with an overridden GetReward) is a user plugin
    return typeid(*f) == typeid(EventReward) || typeid(*f) == typeid(VelocityPlayerToBallReward) || typeid(*f) == typeid(VelocityBallToGoalReward) ||
           typeid(*f) == typeid(FaceBallReward) || typeid(*f) == typeid(VelocityReward) || typeid(*f) == typeid(SaveBoostReward) || typeid(*f) == typeid(TouchBallReward);
}

// Built-in plugin objects -> rlg_engine_cfg ("recognises built-in plugin classes by dynamic_cast"); everything else
// is marked for the host-plugin path.
inline PluginPlan PlanFromMatch(const Match& m, int tickSkip, int numArenas, int device = 0, uint64_t seed = 123) {
    PluginPlan plan;
    rlg_engine_cfg& c = plan.cfg;
    rlg_engine_cfg_default(&c);
    c.num_arenas = numArenas; c.team_size = m.teamSize; c.spawn_opponents = m.spawnOpponents ? 1 : 0; c.tick_skip = tickSkip;
    c.device = device; c.seed = seed;
    // obs
    if (m.obsBuilder && typeid(*m.obsBuilder) == typeid(DefaultOBS)) c.obs_kind = RLG_OBS_DEFAULT;
    else if (m.obsBuilder && typeid(*m.obsBuilder) == typeid(DefaultOBSPadded)) { c.obs_kind = RLG_OBS_PADDED; c.obs_max_players = static_cast<DefaultOBSPadded*>(m.obsBuilder)->maxPlayers; }
    else if (m.obsBuilder) plan.obsOnHost = true;  // the ring keeps the DefaultOBS row width unless the probe below says otherwise
    else throw std::runtime_error("RLGB200: Match has no OBSBuilder");
    // action parser
    if (!m.actionParser) throw std::runtime_error("RLGB200: Match has no ActionParser");
    if (typeid(*m.actionParser) != typeid(DiscreteAction)) {
        // a user ActionParser (ActionParser.h:11-14): every index through ParseActions once, on two different probe states
        const int n = m.actionParser->GetActionAmount();
        if (n < 1 || n > RLG_MAX_ACTIONS) throw std::runtime_error("RLGB200: ActionParser::GetActionAmount() must be in [1, " + std::to_string(RLG_MAX_ACTIONS) + "]");
        const int P = m.playerAmount > 0 ? m.playerAmount : 2;
        GameState probe[2];
        for (int k = 0; k < 2; k++) {
            probe[k].players.resize(P);
            for (int p = 0; p < P; p++) {
                probe[k].players[p].carId = p + 1;
                probe[k].players[p].team = (Team)(p & 1);
                probe[k].players[p].phys.pos = Vec(k ? 1234.f : -800.f, k ? -2500.f : 300.f, k ? 600.f : 17.f);
                probe[k].players[p].phys.vel = Vec(k ? 900.f : 0.f, k ? -300.f : 0.f, 0.f);
                probe[k].players[p].boostFraction = k ? 0.f : 1.f;
                probe[k].players[p].carState.isOnGround = k == 0;
            }
            probe[k].ball.pos = Vec(k ? -2000.f : 0.f, k ? 4000.f : 0.f, k ? 1500.f : 93.f);
        }
        plan.actionTable.resize((size_t)n * 8);
        for (int i = 0; i < n; i++) {
            for (int k = 0; k < 2; k++) {
                ActionSet set = m.actionParser->ParseActions(IList((size_t)P, i), probe[k]);
                if ((int)set.size() != P) throw std::runtime_error("RLGB200: ActionParser::ParseActions must return one Action per player");
                for (int p = 0; p < P; p++) {
                    const Action& a = set[p];
                    const float row[8] = {a.throttle, a.steer, a.pitch, a.yaw, a.roll, a.jump, a.boost, a.handbrake};
                    if (k == 0 && p == 0) memcpy(&plan.actionTable[(size_t)i * 8], row, sizeof(row));
                    else if (memcmp(&plan.actionTable[(size_t)i * 8], row, sizeof(row)) != 0)
                        throw std::runtime_error("RLGB200: this ActionParser maps an index to different Actions for different players / game states; "
                                                 "only state-independent parsers (index -> Action tables) run on the device engine");
                }
            }
        }
    }
    // rewards
    RewardFunction* rf = m.rewardFn;
    c.zero_sum = 0;
    c.num_reward_terms = 0;
    bool fused = rf != nullptr;
    if (rf && typeid(*rf) == typeid(ZeroSumReward)) {
        auto* z = static_cast<ZeroSumReward*>(rf);
        c.zero_sum = 1; c.team_spirit = z->teamSpirit; c.opponent_scale = z->opponentScale; rf = z->childFunc;
    }
    std::vector<std::pair<RewardFunction*, float>> terms;
    if (rf && typeid(*rf) == typeid(CombinedReward)) {
        auto* cr = static_cast<CombinedReward*>(rf);
        for (size_t i = 0; i < cr->rewardFuncs.size(); i++) terms.push_back({cr->rewardFuncs[i], cr->rewardWeights[i]});
    } else if (rf) terms.push_back({rf, 1.f});
    if (terms.size() > RLG_MAX_REWARD_TERMS) fused = false;
    for (size_t i = 0; fused && i < terms.size(); i++) fused = BuiltinRewardTerm(terms[i].first, terms[i].second, c.reward_terms[i]);
    if (fused) c.num_reward_terms = (int32_t)terms.size();
    else { plan.rewardOnHost = true; c.num_reward_terms = 0; c.zero_sum = 0; }
    // terminals
    c.no_touch_max_steps = 0; c.goal_score_terminal = 0;
    for (auto* tc : m.terminalConditions) {
        if (typeid(*tc) == typeid(NoTouchCondition)) c.no_touch_max_steps = (int32_t)static_cast<NoTouchCondition*>(tc)->maxSteps;
        else if (typeid(*tc) == typeid(GoalScoreCondition)) c.goal_score_terminal = 1;
        else plan.hostTerminals++;
    }
    // state setter
    if (dynamic_cast<KickoffState*>(m.stateSetter)) c.state_setter = RLG_SETTER_KICKOFF;
    else if (auto* rs = dynamic_cast<RandomState*>(m.stateSetter)) {
        c.state_setter = RLG_SETTER_RANDOM; c.rand_ball_speed = rs->randBallSpeed; c.rand_car_speed = rs->randCarSpeed; c.cars_on_ground = rs->carsOnGround;
    } else c.state_setter = RLG_SETTER_HOST;  // user StateSetter::ResetState(Arena*) runs on the host
    return plan;
}
inline rlg_engine_cfg CfgFromMatch(const Match& m, int tickSkip, int numArenas, int device = 0, uint64_t seed = 123) {
    return PlanFromMatch(m, tickSkip, numArenas, device, seed).cfg;
}

class Engine {  // RAII over rlg_engine
public:
    rlg_engine* h = nullptr;
    rlg_engine_cfg cfg;
    explicit Engine(const rlg_engine_cfg& c) : cfg(c) { Check(rlg_engine_create(&cfg, &h)); }
    ~Engine() { rlg_engine_destroy(h); }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    void LoadMeshes(const std::vector<std::string>& blobs) {
        std::vector<const void*> p; std::vector<size_t> s;
        for (auto& b : blobs) { p.push_back(b.data()); s.push_back(b.size()); }
        Check(rlg_engine_load_meshes(h, p.data(), s.data(), (int)p.size()));
    }
    int NumArenas() const { return rlg_engine_num_arenas(h); }
    int NumPlayers() const { return rlg_engine_num_players(h); }
    int ObsSize() const { return rlg_engine_obs_size(h); }
};

// Page-locked host array (rlg_host_alloc), falls back to the heap when pinning fails.
template <typename T>
struct HostArray {
    T* p = nullptr;
    size_t n = 0;
    bool pinned = false;
    HostArray() = default;
    HostArray(const HostArray&) = delete;
    HostArray& operator=(const HostArray&) = delete;
    void Resize(size_t count) {
        if (count <= n) return;
        Free();
        p = static_cast<T*>(rlg_host_alloc(count * sizeof(T)));
        pinned = p != nullptr;
        if (!p) p = static_cast<T*>(malloc(count * sizeof(T)));
        if (!p) throw std::runtime_error("RLGB200: out of host memory");
        n = count;
    }
    void Free() { if (p) { if (pinned) rlg_host_free(p); else free(p); } p = nullptr; n = 0; }
    ~HostArray() { Free(); }
    T& operator[](size_t i) { return p[i]; }
};

// rlg_car_state + rlg_gym_player -> PlayerData (PlayerData::UpdateFromCar, PlayerData.cpp:4-33)
inline void FillPlayer(PlayerData& pd, const rlg_car_state& c, const rlg_gym_player& g) {
    pd.carId = (uint32_t)c.car_id;
    pd.team = c.team ? Team::ORANGE : Team::BLUE;
    CarState& s = pd.carState;
    s.pos = Vec(c.pos[0], c.pos[1], c.pos[2]);
    s.rotMat.forward = Vec(c.rot_forward[0], c.rot_forward[1], c.rot_forward[2]);
    s.rotMat.right = Vec(c.rot_right[0], c.rot_right[1], c.rot_right[2]);
    s.rotMat.up = Vec(c.rot_up[0], c.rot_up[1], c.rot_up[2]);
    s.vel = Vec(c.vel[0], c.vel[1], c.vel[2]);
    s.angVel = Vec(c.ang_vel[0], c.ang_vel[1], c.ang_vel[2]);
    s.isOnGround = c.is_on_ground != 0;
    for (int k = 0; k < 4; k++) s.wheelsWithContact[k] = c.wheels_with_contact[k] != 0;
    s.hasJumped = c.has_jumped != 0; s.hasDoubleJumped = c.has_double_jumped != 0; s.hasFlipped = c.has_flipped != 0;
    s.flipRelTorque = Vec(c.flip_rel_torque[0], c.flip_rel_torque[1], c.flip_rel_torque[2]);
    s.jumpTime = c.jump_time; s.flipTime = c.flip_time; s.isFlipping = c.is_flipping != 0; s.isJumping = c.is_jumping != 0;
    s.airTime = c.air_time; s.airTimeSinceJump = c.air_time_since_jump; s.boost = c.boost; s.timeSpentBoosting = c.time_spent_boosting;
    s.isSupersonic = c.is_supersonic != 0; s.supersonicTime = c.supersonic_time; s.handbrakeVal = c.handbrake_val;
    s.isAutoFlipping = c.is_auto_flipping != 0; s.autoFlipTimer = c.auto_flip_timer; s.autoFlipTorqueScale = c.auto_flip_torque_scale;
    s.worldContact.hasContact = c.world_contact_has != 0;
    s.worldContact.contactNormal = Vec(c.world_contact_normal[0], c.world_contact_normal[1], c.world_contact_normal[2]);
    s.carContact.otherCarID = (uint32_t)c.car_contact_other_id; s.carContact.cooldownTimer = c.car_contact_cooldown;
    s.isDemoed = c.is_demoed != 0; s.demoRespawnTimer = c.demo_respawn_timer;
    s.ballHitInfo.isValid = c.hit_valid != 0;
    s.ballHitInfo.relativePosOnBall = Vec(c.hit_rel_pos_on_ball[0], c.hit_rel_pos_on_ball[1], c.hit_rel_pos_on_ball[2]);
    s.ballHitInfo.ballPos = Vec(c.hit_ball_pos[0], c.hit_ball_pos[1], c.hit_ball_pos[2]);
    s.ballHitInfo.extraHitVel = Vec(c.hit_extra_vel[0], c.hit_extra_vel[1], c.hit_extra_vel[2]);
    s.ballHitInfo.tickCountWhenHit = (uint64_t)c.hit_tick; s.ballHitInfo.tickCountWhenExtraImpulseApplied = (uint64_t)c.hit_extra_tick;
    s.lastControls = CarControls{c.last_controls.throttle, c.last_controls.steer, c.last_controls.pitch, c.last_controls.yaw, c.last_controls.roll,
                                 c.last_controls.jump != 0, c.last_controls.boost != 0, c.last_controls.handbrake != 0};
    pd.phys = PhysObj(s);
    pd.physInv = pd.phys.Invert();
    pd.matchGoals = g.match_goals; pd.matchSaves = g.match_saves; pd.matchAssists = g.match_assists; pd.matchShots = g.match_shots;
    pd.matchShotPasses = g.match_shot_passes; pd.matchBumps = g.match_bumps; pd.matchDemos = g.match_demos; pd.boostPickups = g.boost_pickups;
    pd.ballTouchedStep = g.ball_touched_step != 0; pd.ballTouchedTick = g.ball_touched_tick != 0;
    pd.hasJump = !s.hasJumped;
    pd.hasFlip = !s.hasDoubleJumped && !s.hasFlipped && s.airTimeSinceJump < 1.25f;  // RLConst::DOUBLEJUMP_MAX_DELAY
    pd.boostFraction = s.boost / 100;
}
// GameState::UpdateFromArena (GameState.cpp:52-104) from one arena's export
inline void FillGameState(GameState& st, const rlg_gym_state& g, const rlg_ball_state& b, const rlg_car_state* cars, const rlg_gym_player* players, int P) {
    const int64_t dt = g.tick_count - (int64_t)st.lastTickCount;
    st.deltaTime = (float)(dt > 0 ? dt : 0) * (1 / 120.f);
    st.ballState.pos = Vec(b.pos[0], b.pos[1], b.pos[2]); st.ballState.vel = Vec(b.vel[0], b.vel[1], b.vel[2]);
    st.ballState.angVel = Vec(b.ang_vel[0], b.ang_vel[1], b.ang_vel[2]);
    st.ball = PhysObj(st.ballState);
    st.ballInv = st.ball.Invert();
    st.players.resize(P);
    for (int p = 0; p < P; p++) FillPlayer(st.players[p], cars[p], players[p]);
    st.lastTouchCarID = g.last_touch_car_id;
    for (int i = 0; i < RLG_NUM_PADS; i++) {
        const int inv = RLG_NUM_PADS - i - 1;
        st.boostPads[i] = g.pad_active[i] != 0; st.boostPadsInv[i] = g.pad_active[inv] != 0;
        st.boostPadTimers[i] = g.pad_cooldown[i]; st.boostPadTimersInv[i] = g.pad_cooldown[inv];
    }
    st.scoreLine[0] = g.score_line[0]; st.scoreLine[1] = g.score_line[1];
    st.lastTickCount = (uint64_t)g.tick_count;
}
inline void ParallelFor(int n, int threads, const std::function<void(int, int)>& body) {  // body(begin, end)
    threads = std::max(1, std::min(threads, n));
    if (threads == 1) { body(0, n); return; }
    std::vector<std::thread> pool;
    std::vector<std::exception_ptr> errs(threads);
    const int chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t] {
            try { body(std::min(n, t * chunk), std::min(n, (t + 1) * chunk)); } catch (...) { errs[t] = std::current_exception(); }
        });
    for (auto& th : pool) th.join();
    for (auto& e : errs) if (e) std::rethrow_exception(e);
}

// Runs a user StateSetter on the host for the arenas in `ids` and uploads the result (StateSetter::ResetState(Arena*),
// G/Envs/Match.cpp:54-70): cars/ball the setter did not touch keep the default kickoff-spawn state.  `setterOf(i)` gives the
// setter of ids[i] (one Match per arena on the host-plugin path, one shared probe otherwise).
inline void RunHostStateSetter(Engine& e, const std::function<StateSetter*(int)>& setterOf, const std::vector<int32_t>& ids, float* obsOut = nullptr) {
    const int P = e.NumPlayers(), n = (int)ids.size();
    if (n == 0) return;
    static thread_local std::vector<rlg_car_state> cars;
    static thread_local std::vector<rlg_ball_state> balls;
    static thread_local std::vector<rlg_pad_state> pads;
    static thread_local std::vector<int64_t> ticks;
    static thread_local std::vector<uint8_t> mask;
    cars.assign((size_t)n * P, rlg_car_state{}); balls.assign(n, rlg_ball_state{}); pads.assign((size_t)n * RLG_NUM_PADS, rlg_pad_state{}); ticks.assign(n, -1);
    for (int i = 0; i < n; i++) {
        std::vector<RocketSim::Car> carObjs(P);
        RocketSim::Ball ball;
        RocketSim::Arena arena;
        for (int c = 0; c < P; c++) {
            carObjs[c].id = c + 1;
            carObjs[c].team = (e.cfg.spawn_opponents && (c & 1)) ? Team::ORANGE : Team::BLUE;
            arena._cars.push_back(&carObjs[c]);
        }
        arena.ball = &ball;
        setterOf(i)->ResetState(&arena);
        for (int c = 0; c < P; c++) {
            rlg_car_state& o = cars[(size_t)i * P + c];
            const CarState& s = carObjs[c].state;
            o.car_id = c + 1; o.team = (int32_t)carObjs[c].team;
            o.pos[0] = s.pos.x; o.pos[1] = s.pos.y; o.pos[2] = s.pos.z;
            const Vec* cols[3] = {&s.rotMat.forward, &s.rotMat.right, &s.rotMat.up};
            float* dst[3] = {o.rot_forward, o.rot_right, o.rot_up};
            for (int k = 0; k < 3; k++) { dst[k][0] = cols[k]->x; dst[k][1] = cols[k]->y; dst[k][2] = cols[k]->z; }
            o.vel[0] = s.vel.x; o.vel[1] = s.vel.y; o.vel[2] = s.vel.z;
            o.ang_vel[0] = s.angVel.x; o.ang_vel[1] = s.angVel.y; o.ang_vel[2] = s.angVel.z;
            o.is_on_ground = s.isOnGround; o.has_jumped = s.hasJumped; o.has_double_jumped = s.hasDoubleJumped; o.has_flipped = s.hasFlipped;
            o.is_flipping = s.isFlipping; o.is_jumping = s.isJumping; o.jump_time = s.jumpTime; o.flip_time = s.flipTime;
            o.air_time = s.airTime; o.air_time_since_jump = s.airTimeSinceJump;
            o.boost = s.boost; o.is_demoed = s.isDemoed; o.demo_respawn_timer = s.demoRespawnTimer;
            o.is_supersonic = s.isSupersonic; o.supersonic_time = s.supersonicTime; o.handbrake_val = s.handbrakeVal;
            o.hit_tick = -1; o.hit_extra_tick = -1;
        }
        balls[i].pos[0] = ball.state.pos.x; balls[i].pos[1] = ball.state.pos.y; balls[i].pos[2] = ball.state.pos.z;
        balls[i].vel[0] = ball.state.vel.x; balls[i].vel[1] = ball.state.vel.y; balls[i].vel[2] = ball.state.vel.z;
        balls[i].ang_vel[0] = ball.state.angVel.x; balls[i].ang_vel[1] = ball.state.angVel.y; balls[i].ang_vel[2] = ball.state.angVel.z;
        for (int p = 0; p < RLG_NUM_PADS; p++) { auto& ps = pads[(size_t)i * RLG_NUM_PADS + p]; ps.is_active = 1; ps.cooldown = 0; ps.prev_locked_car_id = 0; }  // Match.cpp:66-67
    }
    Check(rlg_engine_set_state(e.h, ids.data(), n, cars.data(), balls.data(), pads.data(), ticks.data()));
    mask.assign(e.NumArenas(), 0);
    for (int32_t id : ids) mask[id] = 1;
    if (obsOut) Check(rlg_engine_reset_current_to(e.h, mask.data(), obsOut, nullptr));
    else Check(rlg_engine_reset_current(e.h, mask.data(), nullptr));
    Check(rlg_engine_sync(e.h));
}
inline void RunHostStateSetter(Engine& e, StateSetter& setter, const std::vector<int32_t>& ids, float* obsOut = nullptr) {
    RunHostStateSetter(e, [&](int) { return &setter; }, ids, obsOut);
}
}  // namespace RLGB200

// ---------------------------------------------------------------------------------------------------------------------
namespace RLGPC {
using RLGSC::IList; using RLGSC::FList; using RLGSC::FList2;

struct Report {  // P/public/RLGymPPO_CPP/Util/Report.h
    typedef double Val;
    std::map<std::string, Val> data;
    Val& operator[](const std::string& key) { return data[key]; }
    Val operator[](const std::string& key) const { return data.at(key); }
    bool Has(const std::string& key) const { return data.find(key) != data.end(); }
    void Accum(const std::string& key, Val v) { if (Has(key)) data[key] += v; else data[key] = v; }
    void AccumAvg(const std::string& key, Val v) { Accum(key + "_avg_total", v); Accum(key + "_avg_count", 1); }
    Val GetAvg(const std::string& key) const {
        const Val total = data.at(key + "_avg_total"), count = data.at(key + "_avg_count");
        return count > 0 ? total / count : 0;
    }
    std::string SingleToString(const std::string& key, bool = false) const {
        std::ostringstream o;
        const Val v = (*this)[key];
        o << key << ": ";
        if ((std::fabs(v) < 1e-3 && v != 0) || std::fabs(v) >= 1e11) o << std::scientific << v;
        else if (v == (Val)(int64_t)v) o << (int64_t)v;
        else o << std::fixed << std::setprecision(4) << v;
        return o.str();
    }
    std::string ToString(bool digitCommas = false, const std::string& prefix = {}) const {
        std::ostringstream o;
        for (auto& kv : data) o << prefix << SingleToString(kv.first, digitCommas) << std::endl;
        return o.str();
    }
    void Clear() { data.clear(); }
    Report operator+(const Report& other) const { Report r = *this; r.data.insert(other.data.begin(), other.data.end()); return r; }
    Report& operator+=(const Report& other) { *this = *this + other; return *this; }
};
struct AvgTracker {  // Util/AvgTracker.h
    float total = 0;
    uint64_t count = 0;
    float Get() const { return count > 0 ? total / count : NAN; }
    void Add(float v) { if (!std::isnan(v)) { total += v; count++; } }
    void Add(float totalVal, uint64_t n) { if (!std::isnan(totalVal)) { total += totalVal; count += n; } }
    AvgTracker& operator+=(float v) { Add(v); return *this; }
    AvgTracker& operator+=(const AvgTracker& o) { Add(o.total, o.count); return *this; }
    void Reset() { total = 0; count = 0; }
};
struct Timer {  // Util/Timer.h
    std::chrono::steady_clock::time_point start = std::chrono::steady_clock::now();
    double Elapsed() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count(); }
    void Reset() { start = std::chrono::steady_clock::now(); }
};
struct WelfordRunningStat {  // Util/WelfordRunningStat.h:36-83
    std::vector<double> runningMean, runningVariance;
    int64_t count = 0, shape = 0;
    WelfordRunningStat() = default;
    explicit WelfordRunningStat(int shape_) : runningMean(shape_), runningVariance(shape_), shape(shape_) {}
    void Update(const FList& sample) {
        const int64_t cur = count++;
        for (int i = 0; i < shape; i++) {
            const double delta = sample[i] - runningMean[i], deltaN = delta / count;
            runningMean[i] += deltaN;
            runningVariance[i] += delta * deltaN * cur;
        }
    }
    void Increment(const FList& samples, int num) { for (int i = 0; i < num; i++) Update(FList{samples[i]}); }
    void Increment(const FList2& samples, int num) { for (int i = 0; i < num; i++) Update(samples[i]); }
    void Reset() { *this = WelfordRunningStat((int)shape); }
    FList Mean() const { FList m(shape, count < 2 ? 0.f : 0.f); if (count >= 2) for (int i = 0; i < shape; i++) m[i] = (float)runningMean[i]; return m; }
    FList GetSTD() const {
        FList s(shape, 1.f);
        if (count < 2) return s;
        for (int i = 0; i < shape; i++) { double v = runningVariance[i] / (count - 1); if (v == 0) v = 1; s[i] = (float)std::sqrt(v); }
        return s;
    }
};

struct PPOLearnerConfig {  // P/public/RLGymPPO_CPP/PPO/PPOLearnerConfig.h:6-32
    IList policyLayerSizes = {256, 256, 256};
    IList criticLayerSizes = {256, 256, 256};
    int64_t batchSize = 50 * 1000;
    int epochs = 10;
    float policyLR = 3e-4f, criticLR = 3e-4f, entCoef = 0.005f, clipRange = 0.2f;
    int64_t miniBatchSize = 0;
    bool autocastLearn = false, halfPrecModels = false;
    float policyTemperature = 1;
    bool measureGradientNoise = false;
    int gradientNoiseUpdateInterval = 10;
    float gradientNoiseAvgDecay = 0.9925f;
};
struct SkillTrackerConfig {  // Util/SkillTrackerConfig.h (the ELO pool is wired in the Python host: rlgymppo_cpp_b200/skill_tracker.py)
    bool enabled = false;
    int numEnvs = 4;
    float simTime = 60;
    int updateInterval = 4;
    int64_t timestepsPerVersion = 50 * 1000 * 1000;
    int maxVersions = 4, numThreads = 8;
    bool perModeRatings = true, loadOldVersionsFromCheckpoints = true, startWithVersion = true, kickoffStatesOnly = true;
    float ratingInc = 5, initialRating = 1000;
};
struct LearnerConfig {  // P/public/RLGymPPO_CPP/LearnerConfig.h:14-81
    int numThreads = 8, numGamesPerThread = 16, minInferenceSize = 80;
    bool renderMode = false;
    float renderTimeScale = 1.5f;
    bool renderDuringTraining = false;
    uint64_t timestepLimit = 0;
    int64_t expBufferSize = 100 * 1000, timestepsPerIteration = 50 * 1000;
    bool standardizeReturns = true, standardizeOBS = false;
    int maxReturnsPerStatsInc = 150, stepsPerObsStatsInc = 5;
    bool deterministic = false, collectionDuringLearn = false;
    PPOLearnerConfig ppo = {};
    float gaeLambda = 0.95f, gaeGamma = 0.99f, rewardClipRange = 10;
    std::string checkpointLoadFolder = "checkpoints", checkpointSaveFolder = "checkpoints";
    bool saveFolderAddUnixTimestamp = false;
    int64_t timestepsPerSave = 500 * 1000;
    int randomSeed = 123, checkpointsToKeep = 5;
    bool sendMetrics = true;
    std::string metricsProjectName = "rlgymppo-cpp", metricsGroupName = "unnamed-runs", metricsRunName = "rlgymppo-cpp-run";
    SkillTrackerConfig skillTrackerConfig = {};
    int deviceIndex = 0;  // (not in the reference: LearnerConfig::deviceType picks CPU/CUDA there) CUDA ordinal of this process
};
struct EnvCreateResult {  // GameInst.h:10-13
    RLGSC::Match* match;
    RLGSC::Gym* gym;
};
typedef std::function<EnvCreateResult()> EnvCreateFn;
class GameInst;
typedef std::function<void(GameInst*, const RLGSC::Gym::StepResult&, Report&)> StepCallback;  // GameInst.h:7

// GameInst.h:16-60.  One per arena on the host-plugin path (it owns that arena's Match / Gym with the user's plugin objects); the
// fused path keeps a single probe instance.  Step / Start happen batched inside ThreadAgentManager::CollectTimesteps.
class GameInst {
public:
    bool isEval = false;
    RLGSC::Gym* gym;
    RLGSC::Match* match;
    FList2 curObs;
    uint64_t totalSteps = 0;
    float curEpRew = 0;
    AvgTracker avgStepRew, avgEpRew;
    Report _metrics = {};
    StepCallback stepCallback = nullptr;
    int arenaIndex = 0;  // (not in the reference) this game's arena in the device pool
    GameInst(RLGSC::Gym* gym_, RLGSC::Match* match_) : gym(gym_), match(match_) {}
    GameInst(const GameInst&) = delete;
    GameInst& operator=(const GameInst&) = delete;
    void ResetMetrics() { avgStepRew.Reset(); avgEpRew.Reset(); _metrics.Clear(); }
    ~GameInst() { delete gym; delete match; }
};

// GameTrajectory.h:5-18 as DEVICE tensors in the reference's concatenated row order (size rows; row i = player n, step t: n * T + t)
struct TrajectoryTensors {
    float* states = nullptr; int64_t* actions = nullptr; float* logProbs = nullptr; float* rewards = nullptr; float* nextStates = nullptr;
    float* dones = nullptr; float* truncateds = nullptr;
};
struct GameTrajectory {
    TrajectoryTensors data;
    size_t size = 0, capacity = 0;
    int obsSize = 0;
};

// Handles with the reference's class names (P/private/RLGymPPO_CPP/PPO/DiscretePolicy.h, ExperienceBuffer.h, torch::Device): the
// networks and the experience FIFO live inside the device learner (rlg_ppo, csrc/ppo.cu).
struct Device { int index = 0; bool is_cpu() const { return false; } bool is_cuda() const { return true; } };
class DiscretePolicy { public: rlg_ppo* ppo = nullptr; float temperature = 1; explicit DiscretePolicy(rlg_ppo* p = nullptr, float t = 1) : ppo(p), temperature(t) {} };
class ExperienceBuffer {
public:
    rlg_ppo* ppo = nullptr;
    explicit ExperienceBuffer(rlg_ppo* p = nullptr) : ppo(p) {}
    int64_t curSize() const { return ppo ? rlg_ppo_buffer_size(ppo) : 0; }
};

// ThreadAgentManager.h:10-60 over ONE device engine: amount x gamesPerAgent arenas; host threads only for user plugins.
class ThreadAgentManager {
public:
    DiscretePolicy* policy; DiscretePolicy* policyHalf;
    ExperienceBuffer* expBuffer;
    bool standardizeOBS, deterministic, blockConcurrentInfer;
    uint64_t maxCollect;
    Device device;
    bool disableCollection = false;
    Timer iterationTimer = {};
    double lastIterationTime = 0;
    // (not in the reference) what the engine needs from LearnerConfig: set these before CreateAgents
    PPOLearnerConfig ppoCfg = {};
    int randomSeed = 123;
    int hostThreads = 8;
    float gaeGamma = 0.99f, gaeLambda = 0.95f, rewardClipRange = 10;

    std::unique_ptr<RLGB200::Engine> engine;
    rlg_collector* collector = nullptr;
    RLGB200::PluginPlan plan;
    std::vector<GameInst*> gameInsts;  // 1 (fused path) or one per arena (host-plugin path)
    StepCallback stepCallback = nullptr;
    int stepsPerCollect = 0;
    bool hostPath = false;

    ThreadAgentManager(DiscretePolicy* policy_, DiscretePolicy* policyHalf_, ExperienceBuffer* expBuffer_, bool standardizeOBS_, bool deterministic_,
                       bool blockConcurrentInfer_, uint64_t maxCollect_, Device device_)
        : policy(policy_), policyHalf(policyHalf_), expBuffer(expBuffer_), standardizeOBS(standardizeOBS_), deterministic(deterministic_),
          blockConcurrentInfer(blockConcurrentInfer_), maxCollect(maxCollect_), device(device_) {
        if (standardizeOBS) throw std::runtime_error("RLGB200: standardizeOBS is not supported (the reference marks it experimental; LearnerConfig.h:41)");
    }
    ThreadAgentManager(const ThreadAgentManager&) = delete;
    ThreadAgentManager& operator=(const ThreadAgentManager&) = delete;
    ~ThreadAgentManager() {
        if (collector) rlg_collector_destroy(collector);
        for (auto* g : gameInsts) delete g;
        FreeTraj();
    }

    void CreateAgents(EnvCreateFn func, int amount, int gamesPerAgent) {
        const int A = amount * gamesPerAgent;
        EnvCreateResult probe = func();  // ThreadAgent.cpp:197-206 makes one Match/Gym per game; the fused path needs one to read the configuration
        gameInsts.push_back(new GameInst(probe.gym, probe.match));
        plan = RLGB200::PlanFromMatch(*probe.match, probe.gym->tickSkip, A, device.index, (uint64_t)randomSeed);
        rlg_engine_cfg& ec = plan.cfg;
        ec.car_preset = probe.gym->carConfig.preset;
        ec.mutators = probe.gym->mutatorConfig.ToC();  // Gym.cpp:43 arena->SetMutatorConfig
        ec.mutators_set = 1;
        engine.reset(new RLGB200::Engine(ec));
        engine->LoadMeshes(RocketSim::CollisionMeshBlobs());
        if (!plan.actionTable.empty())  // a user ActionParser: its table replaces DiscreteAction's, the policy head follows (before the collector)
            RLGB200::Check(rlg_engine_set_action_table(engine->h, plan.actionTable.data(), (int)(plan.actionTable.size() / 8)));
        if (plan.obsOnHost) {  // the ring's row width is the fused builder's: a user builder must produce rows of that width
            hostObsProbe = true;
        }
        const int N = engine->NumArenas() * engine->NumPlayers();
        stepsPerCollect = (int)std::max<int64_t>(1, ((int64_t)maxCollect + N - 1) / N);
        rlg_collector_cfg cc;
        memset(&cc, 0, sizeof(cc));
        if (ppoCfg.policyLayerSizes.size() != ppoCfg.criticLayerSizes.size() || ppoCfg.policyLayerSizes.size() > RLG_MAX_HIDDEN_LAYERS)
            throw std::runtime_error("RLGB200: policy/critic need the same number (<= 4) of hidden layers");
        cc.num_hidden = (int32_t)ppoCfg.policyLayerSizes.size();
        for (int i = 0; i < cc.num_hidden; i++) { cc.policy_hidden[i] = ppoCfg.policyLayerSizes[i]; cc.critic_hidden[i] = ppoCfg.criticLayerSizes[i]; }
        cc.max_steps = stepsPerCollect; cc.seed = (uint64_t)randomSeed; cc.temperature = ppoCfg.policyTemperature; cc.deterministic = deterministic;
        RLGB200::Check(rlg_collector_create(engine->h, &cc, &collector));
        createFn = func;
        UpdateHostPath();
    }
    void SetStepCallback(StepCallback callback) {  // ThreadAgentManager.h:54
        stepCallback = callback;
        for (auto* g : gameInsts) g->stepCallback = callback;
        if (engine) UpdateHostPath();
    }
    // torch nn.Linear tensors of DiscretePolicy (net 0) / ValueEstimator (net 1) from host memory (the Learner pushes device to device)
    void SetLayer(int net, int layer, const float* W, const float* b, int outDim, int inDim) {
        RLGB200::Check(rlg_collector_set_layer(collector, net, layer, W, b, outDim, inDim));
    }
    void StartAgents() {
        if (started) return;
        started = true;
        ResetArenas(AllIds(), nullptr);
    }
    void StopAgents() { RLGB200::Check(rlg_engine_sync(engine->h)); }

    // Blocks until >= amount player-steps are collected (ThreadAgentManager.cpp:16-80) and returns them in the reference's
    // concatenated order (device tensors owned by the manager, valid until the next call).
    GameTrajectory CollectTimesteps(uint64_t amount) {
        CollectOnly(amount);
        return ExportTrajectory();
    }
    // The collection half alone (the Learner feeds the device learner straight from the ring and skips the 7-tensor export).
    void CollectOnly(uint64_t amount) {
        if (!started) StartAgents();
        const uint64_t N = (uint64_t)engine->NumArenas() * engine->NumPlayers();
        int steps = (int)std::max<uint64_t>(1, (amount + N - 1) / N);
        if (steps > stepsPerCollect) throw std::runtime_error("CollectTimesteps: amount exceeds maxCollect");
        Timer t;
        RLGB200::Check(rlg_collector_collect(collector, steps, nullptr));
        RLGB200::Check(rlg_engine_sync(engine->h));
        lastIterationTime = t.Elapsed();
    }
    GameTrajectory ExportTrajectory() {
        rlg_traj_view v;
        RLGB200::Check(rlg_collector_view(collector, &v));
        const size_t rows = (size_t)v.T * v.N;
        if (rows > trajCap) {
            FreeTraj();
            traj.states = (float*)DevAlloc(rows * v.obs_size * 4); traj.nextStates = (float*)DevAlloc(rows * v.obs_size * 4);
            traj.actions = (int64_t*)DevAlloc(rows * 8); traj.logProbs = (float*)DevAlloc(rows * 4); traj.rewards = (float*)DevAlloc(rows * 4);
            traj.dones = (float*)DevAlloc(rows * 4); traj.truncateds = (float*)DevAlloc(rows * 4);
            trajCap = rows;
        }
        RLGB200::Check(rlg_collector_export(collector, traj.states, traj.actions, traj.logProbs, traj.rewards, traj.nextStates, traj.dones, traj.truncateds,
                                            nullptr, nullptr, nullptr));
        RLGB200::Check(rlg_engine_sync(engine->h));
        GameTrajectory out;
        out.data = traj; out.size = rows; out.capacity = trajCap; out.obsSize = v.obs_size;
        return out;
    }
    rlg_traj_view View() { rlg_traj_view v; RLGB200::Check(rlg_collector_view(collector, &v)); return v; }

    void GetMetrics(Report& report) {  // ThreadAgentManager.cpp:82-117
        double sm = 0, im = 0; int32_t sn = 0, in = 0;
        RLGB200::Check(rlg_collector_kernel_times(collector, &sm, &sn, &im, &in));
        if (hostPath) {
            AvgTracker step, ep;
            for (auto* g : gameInsts) { step += g->avgStepRew; ep += g->avgEpRew; }
            report["Average Step Reward"] = step.Get();
            report["Average Episode Reward"] = ep.Get();
        } else {
            rlg_metrics_host m;
            RLGB200::Check(rlg_engine_metrics(engine->h, &m));
            report["Average Step Reward"] = m.avg_step_reward;
            report["Average Episode Reward"] = m.avg_episode_reward;
        }
        report["Env Step Time"] = sm * 1e-3;
        report["Policy Infer Time"] = im * 1e-3;
    }
    void ResetMetrics() {
        RLGB200::Check(rlg_engine_reset_metrics(engine->h));
        for (auto* g : gameInsts) g->ResetMetrics();
    }

private:
    EnvCreateFn createFn;
    bool started = false, hostObsProbe = false;
    TrajectoryTensors traj;
    size_t trajCap = 0;
    std::vector<void*> devAllocs;
    // host-plugin staging
    RLGB200::HostArray<rlg_car_state> hCars; RLGB200::HostArray<rlg_ball_state> hBalls; RLGB200::HostArray<rlg_gym_state> hGym;
    RLGB200::HostArray<rlg_gym_player> hPlayers; RLGB200::HostArray<float> hObs, hRew; RLGB200::HostArray<uint8_t> hDone;
    std::vector<RLGSC::Gym::StepResult> results;
    std::vector<int32_t> doneIds;
    std::vector<uint8_t> resetMask;

    // cudaMalloc without a CUDA dependency in this header: the collector's ring allocator is not exposed, so trajectory tensors
    // come from a tiny device-memory helper of the C ABI
    void* DevAlloc(size_t bytes) {
        void* p = rlg_device_alloc(engine->h, bytes);
        if (!p) throw std::runtime_error(std::string("RLGB200: ") + rlg_last_error());
        devAllocs.push_back(p);
        return p;
    }
    void FreeTraj() {
        for (void* p : devAllocs) rlg_device_free(engine ? engine->h : nullptr, p);
        devAllocs.clear();
        traj = TrajectoryTensors();
        trajCap = 0;
    }
    std::vector<int32_t> AllIds() const {
        std::vector<int32_t> ids(engine->NumArenas());
        for (size_t i = 0; i < ids.size(); i++) ids[i] = (int32_t)i;
        return ids;
    }
    RLGSC::Match* MatchOf(int arena) { return gameInsts.size() > 1 ? gameInsts[arena]->match : gameInsts[0]->match; }

    // the host-plugin path is on when any stage is user-defined or a StepCallback is installed; it needs one GameInst per arena
    void UpdateHostPath() {
        const bool want = plan.AnyHost() || (bool)stepCallback;
        if (want && (int)gameInsts.size() < engine->NumArenas()) {
            for (int a = (int)gameInsts.size(); a < engine->NumArenas(); a++) {
                EnvCreateResult r = createFn();
                gameInsts.push_back(new GameInst(r.gym, r.match));
            }
            for (int a = 0; a < (int)gameInsts.size(); a++) { gameInsts[a]->arenaIndex = a; gameInsts[a]->stepCallback = stepCallback; }
        }
        hostPath = want;
        RLGB200::Check(rlg_collector_set_step_hook(collector, want ? &ThreadAgentManager::StepHook : nullptr, this));
        RLGB200::Check(rlg_collector_set_reset_hook(collector, (!want && plan.cfg.state_setter == RLG_SETTER_HOST) ? &ThreadAgentManager::ResetHook : nullptr, this));
    }
    static void ResetHook(void* user, const int32_t* ids, int n, float* obsOut) {
        auto* self = static_cast<ThreadAgentManager*>(user);
        self->ResetArenas(std::vector<int32_t>(ids, ids + n), obsOut);
    }
    static int StepHook(void* user, int t, const int32_t* actions, float* obsNext, float* reward, uint8_t* done) {
        try {
            static_cast<ThreadAgentManager*>(user)->HostStep(t, actions, obsNext, reward, done);
            return RLG_OK;
        } catch (std::exception& ex) {
            rlg_set_last_error(ex.what());
            return RLG_ERR_STATE;
        }
    }

    // Gym::Reset for the given arenas (G/Gym.cpp:58-66): state setter (device or host) + Match::EpisodeReset + obs; on the host-plugin
    // path the plugins' Reset virtuals run on the reset GameState and a user OBSBuilder builds the reset observations.
    void ResetArenas(const std::vector<int32_t>& ids, float* obsOut) {
        const int A = engine->NumArenas(), P = engine->NumPlayers(), O = engine->ObsSize(), n = (int)ids.size();
        if (n == 0) return;
        if (plan.cfg.state_setter == RLG_SETTER_HOST) {
            RLGB200::RunHostStateSetter(*engine, [&](int i) { return MatchOf(ids[i])->stateSetter; }, ids, obsOut);
        } else {
            resetMask.assign(A, 0);
            for (int32_t id : ids) resetMask[id] = 1;
            if (obsOut) RLGB200::Check(rlg_engine_reset_to(engine->h, resetMask.data(), obsOut, nullptr));
            else RLGB200::Check(rlg_engine_reset(engine->h, resetMask.data(), nullptr));
        }
        if (!hostPath) return;
        EnsureStaging();
        RLGB200::Check(rlg_engine_export_gamestates(engine->h, ids.data(), n, hCars.p, hBalls.p, hGym.p, hPlayers.p));
        std::vector<float> rows(plan.obsOnHost ? (size_t)n * P * O : 0);
        RLGB200::ParallelFor(n, hostThreads, [&](int b, int e) {
            for (int i = b; i < e; i++) {
                GameInst* g = gameInsts[ids[i]];
                RLGSC::GameState st;  // a fresh GameState(arena), as the state setters return (lastTickCount 0)
                RLGB200::FillGameState(st, hGym[i], hBalls[i], hCars.p + (size_t)i * P, hPlayers.p + (size_t)i * P, P);
                RLGSC::Match* m = g->match;
                m->prevActions = RLGSC::ActionSet(P);  // Match::EpisodeReset (Match.cpp:4-10)
                for (auto* c : m->terminalConditions) c->Reset(st);
                if (plan.rewardOnHost) m->rewardFn->Reset(st);
                if (plan.obsOnHost) {
                    m->obsBuilder->Reset(st);
                    m->obsBuilder->PreStep(st);
                    for (int p = 0; p < P; p++) {
                        const FList o = m->obsBuilder->BuildOBS(st.players[p], st, m->prevActions[p]);
                        CheckObsWidth(o, O);
                        memcpy(rows.data() + ((size_t)i * P + p) * O, o.data(), (size_t)O * 4);
                    }
                }
                g->gym->prevState = st;
            }
        });
        if (plan.obsOnHost) {
            float* dst = obsOut;
            if (!dst) RLGB200::Check(rlg_engine_outputs(engine->h, &dst, nullptr, nullptr));
            // contiguous runs of arena ids go up in one copy
            int i = 0;
            while (i < n) {
                int j = i + 1;
                while (j < n && ids[j] == ids[j - 1] + 1) j++;
                RLGB200::Check(rlg_engine_copy_to_device(engine->h, dst + (size_t)ids[i] * P * O, rows.data() + (size_t)i * P * O, (size_t)(j - i) * P * O * 4));
                i = j;
            }
        }
    }
    static void CheckObsWidth(const FList& o, int O) {
        if ((int)o.size() != O)
            throw std::runtime_error("RLGB200: a user OBSBuilder must produce rows of the engine's observation width (" + std::to_string(O) +
                                     " floats for this mode: DefaultOBS / DefaultOBSPadded layout size), got " + std::to_string(o.size()));
    }
    void EnsureStaging() {
        const size_t A = engine->NumArenas(), P = engine->NumPlayers(), O = engine->ObsSize();
        hCars.Resize(A * P); hBalls.Resize(A); hGym.Resize(A); hPlayers.Resize(A * P); hObs.Resize(A * P * O); hRew.Resize(A * P); hDone.Resize(A);
        results.resize(A);
    }

    // GameInst::Step for every arena (GameInst.cpp:7-38) with the Gym::Step pieces that are user-defined done on the host.
    void HostStep(int, const int32_t* actions, float* obsNext, float* reward, uint8_t* done) {
        const int A = engine->NumArenas(), P = engine->NumPlayers(), O = engine->ObsSize();
        EnsureStaging();
        // the snapshot GameStates travel while the GPU runs ticks 1..tickSkip-1
        RLGB200::Check(rlg_engine_export_gamestates_async(engine->h, nullptr, A, hCars.p, hBalls.p, hGym.p, hPlayers.p));
        RLGB200::Check(rlg_engine_step_end(engine->h, actions, nullptr));
        RLGB200::Check(rlg_engine_export_wait(engine->h));
        const bool needObs = !plan.obsOnHost && (bool)stepCallback;  // StepResult::obs of the fused builder
        if (needObs) RLGB200::Check(rlg_engine_copy_to_host(engine->h, hObs.p, obsNext, (size_t)A * P * O * 4));
        if (!plan.rewardOnHost) RLGB200::Check(rlg_engine_copy_to_host(engine->h, hRew.p, reward, (size_t)A * P * 4));
        RLGB200::Check(rlg_engine_copy_to_host(engine->h, hDone.p, done, (size_t)A));
        RLGB200::ParallelFor(A, hostThreads, [&](int b, int e) {
            for (int a = b; a < e; a++) {
                GameInst* g = gameInsts[a];
                RLGSC::Match* m = g->match;
                RLGSC::Gym::StepResult& r = results[a];
                r.state = g->gym->prevState;  // "state = prevState; state.UpdateFromArena(arena)" (Gym.cpp:86-87)
                RLGB200::FillGameState(r.state, hGym[a], hBalls[a], hCars.p + (size_t)a * P, hPlayers.p + (size_t)a * P, P);
                m->prevActions.resize(P);
                for (int p = 0; p < P; p++) {
                    const float* pa = hPlayers[(size_t)a * P + p].prev_action;
                    m->prevActions[p] = RLGSC::Action{pa[0], pa[1], pa[2], pa[3], pa[4], pa[5], pa[6], pa[7]};
                }
                // Match::BuildObservations (Match.cpp:12-25)
                r.obs.resize(P);
                if (plan.obsOnHost) {
                    m->obsBuilder->PreStep(r.state);
                    for (int p = 0; p < P; p++) {
                        r.obs[p] = m->obsBuilder->BuildOBS(r.state.players[p], r.state, m->prevActions[p]);
                        CheckObsWidth(r.obs[p], O);
                        memcpy(hObs.p + ((size_t)a * P + p) * O, r.obs[p].data(), (size_t)O * 4);
                    }
                } else if (needObs) {
                    for (int p = 0; p < P; p++) r.obs[p].assign(hObs.p + ((size_t)a * P + p) * O, hObs.p + ((size_t)a * P + p + 1) * O);
                }
                // Match::IsDone (Match.cpp:35-41): the fused built-in conditions OR the user's
                bool d = hDone[a] != 0;
                if (plan.hostTerminals > 0)
                    for (auto* c : m->terminalConditions)
                        if (typeid(*c) != typeid(RLGSC::NoTouchCondition) && typeid(*c) != typeid(RLGSC::GoalScoreCondition) && c->IsTerminal(r.state)) d = true;
                r.done = d;
                hDone[a] = d ? 1 : 0;
                // Match::GetRewards (Match.cpp:27-33)
                if (plan.rewardOnHost) {
                    m->rewardFn->PreStep(r.state);
                    r.reward = m->rewardFn->GetAllRewards(r.state, m->prevActions, d);
                    for (int p = 0; p < P; p++) hRew[(size_t)a * P + p] = r.reward[p];
                } else {
                    r.reward.assign(hRew.p + (size_t)a * P, hRew.p + (size_t)(a + 1) * P);
                }
                g->gym->prevState = r.state;
                g->gym->totalSteps++; g->gym->totalTicks += g->gym->tickSkip;
                // GameInst::Step (GameInst.cpp:13-37)
                float total = 0;
                for (int p = 0; p < P; p++) total += r.reward[p];
                g->avgStepRew.Add(total, (uint64_t)P);
                g->curEpRew += total / P;
                if (g->stepCallback) g->stepCallback(g, r, g->_metrics);
                if (d) { g->avgEpRew += g->curEpRew; g->curEpRew = 0; }
                g->totalSteps++;
            }
        });
        if (plan.obsOnHost) RLGB200::Check(rlg_engine_copy_to_device(engine->h, obsNext, hObs.p, (size_t)A * P * O * 4));
        if (plan.rewardOnHost) RLGB200::Check(rlg_engine_copy_to_device(engine->h, reward, hRew.p, (size_t)A * P * 4));
        if (plan.hostTerminals > 0) RLGB200::Check(rlg_engine_copy_to_device(engine->h, done, hDone.p, (size_t)A));
        doneIds.clear();
        for (int a = 0; a < A; a++) if (hDone[a]) doneIds.push_back(a);
        ResetArenas(doneIds, obsNext);  // nextObs = gym->Reset() for finished games (GameInst.cpp:27-28)
    }
};

// PPOLearner (P/private/RLGymPPO_CPP/PPO/PPOLearner.h) over the device learner of csrc/ppo.cu
class PPOLearner {
public:
    rlg_ppo* h = nullptr;
    PPOLearnerConfig config;
    Device device;
    DiscretePolicy* policy = nullptr; DiscretePolicy* policyHalf = nullptr;
    uint64_t cumulativeModelUpdates = 0;
    PPOLearner(int obsSize, int actionAmount, PPOLearnerConfig cfg, Device device_, int64_t expBufferSize, uint64_t seed) : config(cfg), device(device_) {
        if (config.miniBatchSize == 0) config.miniBatchSize = config.batchSize;  // PPOLearner.cpp:19-20
        if (config.batchSize % config.miniBatchSize != 0) throw std::runtime_error("PPOLearner: batchSize must be a multiple of miniBatchSize");  // :22-23
        if (config.policyLayerSizes.size() != config.criticLayerSizes.size() || config.policyLayerSizes.size() > RLG_MAX_HIDDEN_LAYERS)
            throw std::runtime_error("RLGB200: policy/critic need the same number (<= 4) of hidden layers");
        rlg_ppo_cfg c;
        memset(&c, 0, sizeof(c));
        actionAmountValue = actionAmount;
        c.device = device.index; c.obs_size = obsSize; c.num_actions = actionAmount; c.num_hidden = (int32_t)config.policyLayerSizes.size();
        for (int i = 0; i < c.num_hidden; i++) { c.policy_hidden[i] = config.policyLayerSizes[i]; c.critic_hidden[i] = config.criticLayerSizes[i]; }
        c.batch_size = config.batchSize; c.mini_batch_size = config.miniBatchSize; c.epochs = config.epochs;
        c.policy_lr = config.policyLR; c.critic_lr = config.criticLR; c.ent_coef = config.entCoef; c.clip_range = config.clipRange;
        c.temperature = config.policyTemperature; c.exp_buffer_size = expBufferSize; c.seed = seed; c.world = 1;
        RLGB200::Check(rlg_ppo_create(&c, &h));
        RLGB200::Check(rlg_ppo_init_weights(h, seed));  // torch::nn::Linear's default initialisation
        policy = new DiscretePolicy(h, config.policyTemperature);
    }
    PPOLearner(const PPOLearner&) = delete;
    PPOLearner& operator=(const PPOLearner&) = delete;
    ~PPOLearner() { delete policy; rlg_ppo_destroy(h); }
    void Learn(ExperienceBuffer*, Report& report) {  // PPOLearner.cpp:67-349
        rlg_ppo_report r;
        Timer t;
        RLGB200::Check(rlg_ppo_learn(h, &r, nullptr));
        const double total = t.Elapsed();
        if (r.batches == 0) fprintf(stderr, "PPOLearner::Learn(): WARNING: the experience buffer holds fewer rows than batchSize: no optimiser step was taken\n");
        cumulativeModelUpdates += (uint64_t)r.batches;
        report["PPO Batch Consumption Time"] = total / (double)std::max<int64_t>(r.batches, 1);
        report["Cumulative Model Updates"] = (double)cumulativeModelUpdates;
        report["Policy Entropy"] = r.entropy; report["Mean KL Divergence"] = r.kl; report["Mean Ratio"] = r.ratio;
        report["Value Function Loss"] = r.value_loss; report["SB3 Clip Fraction"] = r.clip_fraction;
        report["Policy Update Magnitude"] = r.policy_update_magnitude; report["Value Function Update Magnitude"] = r.critic_update_magnitude;
        report["PPO Learn Time"] = total;
    }
    void UpdateLearningRates(float policyLR, float criticLR) {  // PPOLearner.cpp:504-517
        config.policyLR = policyLR; config.criticLR = criticLR;
        RLGB200::Check(rlg_ppo_set_lr(h, policyLR, criticLR));
        printf("PPOLearner: Updated learning rate to [%e, %e]\n", policyLR, criticLR);
    }
    // PPOLearner::SaveTo / LoadFrom (PPOLearner.cpp:362-502).  File format: this shim has no libtorch, so the networks and Adam
    // moments go to <folder>/PPO_{POLICY,CRITIC}.rlgb (magic, layer shapes, then W, b, exp_avg, exp_avg_sq per layer, float32);
    // the Python host (rlgymppo_cpp_b200/checkpoint.py) reads and writes the reference's TorchScript layout.
    void SaveTo(const std::string& folder) { for (int net = 0; net < 2; net++) NetIO(folder + (net == 0 ? "/PPO_POLICY.rlgb" : "/PPO_CRITIC.rlgb"), net, true); }
    void LoadFrom(const std::string& folder) { for (int net = 0; net < 2; net++) NetIO(folder + (net == 0 ? "/PPO_POLICY.rlgb" : "/PPO_CRITIC.rlgb"), net, false); }
    std::vector<std::pair<int, int>> Dims(int net) const {
        std::vector<std::pair<int, int>> d;
        const IList& hsz = net == 0 ? config.policyLayerSizes : config.criticLayerSizes;
        int in = obsSize_();
        for (int hdim : hsz) { d.push_back({hdim, in}); in = hdim; }
        d.push_back({net == 0 ? actionAmountValue : 1, in});
        return d;
    }
    int obsSizeValue = 0;
    int actionAmountValue = RLG_NUM_ACTIONS;  // the policy head: ActionParser::GetActionAmount()

private:
    int obsSize_() const { return obsSizeValue; }
    void NetIO(const std::string& path, int net, bool save) {
        const auto dims = Dims(net);
        std::fstream f(path, std::ios::binary | (save ? std::ios::out | std::ios::trunc : std::ios::in));
        if (!f.good()) throw std::runtime_error(std::string(save ? "PPOLearner::SaveTo(): cannot write " : "PPOLearner::LoadFrom(): model file does not exist: ") + path);
        const uint32_t magic = 0x42474C52u;  // "RLGB"
        uint32_t hdr[2] = {magic, (uint32_t)dims.size()};
        int64_t steps[2] = {0, 0};
        if (save) {
            RLGB200::Check(rlg_ppo_adam_steps(h, &steps[0], &steps[1], 0));
            f.write((const char*)hdr, sizeof(hdr));
            f.write((const char*)&steps[net], 8);
        } else {
            f.read((char*)hdr, sizeof(hdr));
            if (hdr[0] != magic || hdr[1] != dims.size()) throw std::runtime_error("PPOLearner::LoadFrom(): saved model has a different layer count: " + path);
            int64_t st = 0;
            f.read((char*)&st, 8);
            RLGB200::Check(rlg_ppo_adam_steps(h, &steps[0], &steps[1], 0));
            steps[net] = st;
            RLGB200::Check(rlg_ppo_adam_steps(h, &steps[0], &steps[1], 1));
        }
        for (size_t l = 0; l < dims.size(); l++) {
            int32_t shape[2] = {dims[l].first, dims[l].second};
            if (save) f.write((const char*)shape, 8);
            else {
                int32_t got[2];
                f.read((char*)got, 8);
                if (got[0] != shape[0] || got[1] != shape[1]) throw std::runtime_error("PPOLearner::LoadFrom(): saved model has different size than current model: " + path);
            }
            std::vector<float> W((size_t)shape[0] * shape[1]), b(shape[0]);
            for (int which : {0, 2, 3}) {  // parameters, exp_avg, exp_avg_sq
                if (save) {
                    RLGB200::Check(rlg_ppo_get_layer(h, which, net, (int)l, W.data(), b.data(), shape[0], shape[1]));
                    f.write((const char*)W.data(), W.size() * 4); f.write((const char*)b.data(), b.size() * 4);
                } else {
                    f.read((char*)W.data(), W.size() * 4); f.read((char*)b.data(), b.size() * 4);
                    if (!f.good()) throw std::runtime_error("PPOLearner::LoadFrom(): truncated model file: " + path);
                    RLGB200::Check(rlg_ppo_set_layer(h, which, net, (int)l, W.data(), b.data(), shape[0], shape[1]));
                }
            }
        }
    }
};

class Learner;
typedef std::function<void(Learner*, Report&)> IterationCallback;

// Learner (P/public/RLGymPPO_CPP/Learner.h:14-60, Learner.cpp): same constructor, members and Learn() loop over the device engine
// and the device learner.  One process drives one GPU (LearnerConfig::deviceIndex); multi-GPU data parallelism is the Python
// host's job (rlgymppo_cpp_b200/learner.py under torchrun).
class Learner {
public:
    LearnerConfig config;
    PPOLearner* ppo = nullptr;
    ThreadAgentManager* agentMgr = nullptr;
    ExperienceBuffer* expBuffer = nullptr;
    EnvCreateFn envCreateFn;
    void* metricSender = nullptr;  // wandb lives in the Python host (rlgymppo_cpp_b200/sinks.py); reports are printed here
    void* renderSender = nullptr;
    void* skillTracker = nullptr;  // the ELO pool is wired in the Python host (rlgymppo_cpp_b200/skill_tracker.py)
    int obsSize = 0, actionAmount = 0;
    std::string runID = {};
    uint64_t totalTimesteps = 0, totalEpochs = 0;
    WelfordRunningStat returnStats = WelfordRunningStat(1);
    IterationCallback iterationCallback = nullptr;
    StepCallback stepCallback = nullptr;
    uint64_t maxIterations = 0;  // (not in the reference) 0 = until timestepLimit; tests stop after a few iterations

    Learner(EnvCreateFn envCreateFunc, LearnerConfig cfg) : config(cfg), envCreateFn(envCreateFunc) {
        printf("Learner::Learner():\n");
        if (config.timestepsPerSave == 0) throw std::runtime_error("Learner::Learner(): timestepsPerSave cannot be zero");
        if (config.standardizeOBS) throw std::runtime_error("Learner::Learner(): standardizeOBS is not supported by the device engine");
        if (config.saveFolderAddUnixTimestamp && !config.checkpointSaveFolder.empty())
            config.checkpointSaveFolder += "-" + std::to_string((long long)time(nullptr));  // Learner.cpp:30-31
        if (config.expBufferSize < config.ppo.batchSize)
            throw std::runtime_error("Learner::Learner(): expBufferSize is smaller than ppo.batchSize: no batch would ever be formed");
        Device dev{config.deviceIndex};
        {   // "Creating test environment to determine OBS size and action amount" (Learner.cpp:71-83): read it from the plugin plan
            EnvCreateResult env = envCreateFunc();
            RLGB200::PluginPlan plan = RLGB200::PlanFromMatch(*env.match, env.gym->tickSkip, 1, dev.index, (uint64_t)config.randomSeed);
            const int P = env.match->playerAmount;
            obsSize = plan.cfg.obs_kind == RLG_OBS_PADDED ? 51 + 19 * 2 * plan.cfg.obs_max_players : 51 + 19 * P;
            actionAmount = env.match->actionParser->GetActionAmount();
            delete env.gym; delete env.match;
        }
        printf("\tOBS size: %d, action amount: %d\n", obsSize, actionAmount);
        ppo = new PPOLearner(obsSize, actionAmount, config.ppo, dev, config.expBufferSize, (uint64_t)config.randomSeed);
        ppo->obsSizeValue = obsSize;
        expBuffer = new ExperienceBuffer(ppo->h);
        agentMgr = new ThreadAgentManager(ppo->policy, ppo->policyHalf, expBuffer, config.standardizeOBS, config.deterministic, false,
                                          (uint64_t)config.timestepsPerIteration, dev);
        agentMgr->ppoCfg = config.ppo; agentMgr->randomSeed = config.randomSeed; agentMgr->hostThreads = config.numThreads;
        agentMgr->gaeGamma = config.gaeGamma; agentMgr->gaeLambda = config.gaeLambda; agentMgr->rewardClipRange = config.rewardClipRange;
        printf("\tCreating %d arenas on the device...\n", config.numThreads * config.numGamesPerThread);
        agentMgr->CreateAgents(envCreateFunc, config.numThreads, config.numGamesPerThread);
        if (!config.checkpointLoadFolder.empty()) Load();  // Learner.cpp:130-131
        if (config.sendMetrics) printf("\tNOTE: sendMetrics: the wandb bridge lives in the Python host (rlgymppo_cpp_b200/sinks.py); this host prints its reports.\n");
        RLGB200::Check(rlg_ppo_push_weights(ppo->h, agentMgr->collector, rlg_engine_stream(agentMgr->engine->h)));
    }
    Learner(const Learner&) = delete;
    Learner& operator=(const Learner&) = delete;
    ~Learner() { delete agentMgr; delete expBuffer; delete ppo; }

    void UpdateLearningRates(float policyLR, float criticLR) { config.ppo.policyLR = policyLR; config.ppo.criticLR = criticLR; ppo->UpdateLearningRates(policyLR, criticLR); }
    std::vector<Report> GetAllGameMetrics() {  // Learner.cpp:158-169
        std::vector<Report> out;
        if (agentMgr->hostPath) for (auto* g : agentMgr->gameInsts) out.push_back(g->_metrics);
        return out;
    }

    // Learner::AddNewExperience (Learner.cpp:608-703).  The value predictions and GAE of the reference's body were already taken
    // on the device by the collector (ring + k_gae), so the trajectory argument is only checked for its size.
    void AddNewExperience(GameTrajectory& gameTraj, Report& report) {
        (void)gameTraj;
        AddNewExperienceFromRing(report);
    }
    void AddNewExperienceFromRing(Report& report) {
        rlg_collector* col = agentMgr->collector;
        void* es = rlg_engine_stream(agentMgr->engine->h);
        const float retStd = config.standardizeReturns ? returnStats.GetSTD()[0] : 1.f;
        RLGB200::Check(rlg_collector_gae(col, config.gaeGamma, config.gaeLambda, retStd, config.rewardClipRange, nullptr));
        double means[3] = {0, 0, 0};
        const int nFirst = config.standardizeReturns ? std::min(config.maxReturnsPerStatsInc, 4096) : 0;
        FList first(std::max(nFirst, 1));
        RLGB200::Check(rlg_collector_return_stats(col, means, first.data(), nFirst, nullptr));
        report["Avg Return"] = means[0] / retStd; report["Avg Advantage"] = means[1]; report["Avg Val Target"] = means[2];
        if (config.standardizeReturns) {
            rlg_traj_view v = agentMgr->View();
            returnStats.Increment(first, (int)std::min<int64_t>(nFirst, (int64_t)v.T * v.N));
        }
        RLGB200::Check(rlg_ppo_submit_collector(ppo->h, col, es));
    }

    void Learn() {  // Learner.cpp:436-606
        printf("Learner::Learn():\n\tStarting agents...\n");
        agentMgr->SetStepCallback(stepCallback);
        agentMgr->StartAgents();
        int64_t tsSinceSave = 0;
        uint64_t iterations = 0;
        Timer epochTimer;
        while ((totalTimesteps < config.timestepLimit || config.timestepLimit == 0) && (maxIterations == 0 || iterations < maxIterations)) {
            Report report = {};
            agentMgr->SetStepCallback(stepCallback);
            agentMgr->CollectOnly((uint64_t)config.timestepsPerIteration);
            const rlg_traj_view v = agentMgr->View();
            const double relCollectionTime = epochTimer.Elapsed();
            const uint64_t timestepsCollected = (uint64_t)v.T * v.N;
            totalTimesteps += timestepsCollected;
            iterations++;
            if (config.ppo.policyLR == 0 && config.ppo.criticLR == 0) { printf("\tBoth LRs are set to zero. Skipping consumption!\n"); continue; }
            if (config.deterministic)
                throw std::runtime_error("Learner::Learn(): Cannot run PPO learn iteration when on deterministic mode!\nDeterministic mode is meant for performing, not training. Only collection should occur.");
            AddNewExperienceFromRing(report);
            Timer ppoLearnTimer;
            ppo->Learn(expBuffer, report);
            totalEpochs += config.ppo.epochs;
            RLGB200::Check(rlg_ppo_push_weights(ppo->h, agentMgr->collector, rlg_engine_stream(agentMgr->engine->h)));  // SetNewPolicy
            RLGB200::Check(rlg_engine_sync(agentMgr->engine->h));
            const double relEpochTime = epochTimer.Elapsed();
            epochTimer.Reset();
            agentMgr->GetMetrics(report);
            report["Total Iteration Time"] = relEpochTime; report["Collection Time"] = relCollectionTime;
            report["Consumption Time"] = relEpochTime - relCollectionTime; report["Collect-Consume Overlap Time"] = 0;
            report["Collected Steps/Second"] = (double)(int64_t)(timestepsCollected / std::max(relCollectionTime, 1e-9));
            report["Overall Steps/Second"] = (double)(int64_t)(timestepsCollected / std::max(relEpochTime, 1e-9));
            report["Timesteps Collected"] = (double)timestepsCollected; report["Cumulative Timesteps"] = (double)totalTimesteps;
            if (iterationCallback) iterationCallback(this, report);
            printf("\n============================================\nITERATION COMPLETED:\n\n%s============================================\n\n", report.ToString(true, " ").c_str());
            tsSinceSave += (int64_t)timestepsCollected;
            if (tsSinceSave > config.timestepsPerSave && !config.checkpointSaveFolder.empty()) { Save(); tsSinceSave = 0; }
            agentMgr->ResetMetrics();
        }
        printf("Learner: stopping\n\tStopping agents...\n");
        agentMgr->StopAgents();
    }

    // Learner::Save / Load (Learner.cpp:244-365): <folder>/<timesteps>/RUNNING_STATS.json in the reference's schema + the networks
    void Save() {
        if (config.checkpointSaveFolder.empty()) throw std::runtime_error("Learner::Save(): Cannot save because config.checkpointSaveFolder is not set");
        const std::string dst = config.checkpointSaveFolder + "/" + std::to_string(totalTimesteps);
        MakeDirs(dst);
        printf("Saving to folder %s...\n", dst.c_str());
        SaveStats(dst + "/RUNNING_STATS.json");
        ppo->SaveTo(dst);
        if (config.checkpointsToKeep != -1) {  // Learner.cpp:256-278: drop the lowest-numbered while there are too many
            auto nums = NumberedFolders(config.checkpointLoadFolder);
            if ((int)nums.size() > config.checkpointsToKeep) RemoveTree(config.checkpointLoadFolder + "/" + std::to_string(nums.front()));
        }
    }
    void Load() {
        if (config.checkpointLoadFolder.empty()) throw std::runtime_error("Learner::Load(): Cannot load because config.checkpointLoadFolder is not set");
        auto nums = NumberedFolders(config.checkpointLoadFolder);
        if (nums.empty()) { printf("\tNo checkpoints found in %s, starting a new model.\n", config.checkpointLoadFolder.c_str()); return; }
        const std::string src = config.checkpointLoadFolder + "/" + std::to_string(nums.back());
        printf("\tLoading checkpoint %s...\n", src.c_str());
        LoadStats(src + "/RUNNING_STATS.json");
        ppo->LoadFrom(src);
    }
    void SaveStats(const std::string& path) {
        std::ofstream f(path);
        if (!f.good()) throw std::runtime_error("Learner::SaveStats(): Can't open file at " + path);
        f << std::setprecision(17) << "{\n    \"cumulative_timesteps\": " << totalTimesteps << ",\n    \"cumulative_model_updates\": " << ppo->cumulativeModelUpdates
          << ",\n    \"epoch\": " << totalEpochs << ",\n    \"reward_running_stats\": {\n        \"mean\": [" << returnStats.runningMean[0] << "],\n        \"var\": ["
          << returnStats.runningVariance[0] << "],\n        \"shape\": 1,\n        \"count\": " << returnStats.count << "\n    }";
        if (!runID.empty()) f << ",\n    \"run_id\": \"" << runID << "\"";
        f << "\n}\n";
    }
    void LoadStats(const std::string& path) {
        std::ifstream f(path);
        if (!f.good()) throw std::runtime_error("Learner::LoadStats(): Can't open file at " + path);
        const std::string j((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        auto num = [&](const std::string& key, size_t from = 0) -> double {
            size_t k = j.find("\"" + key + "\"", from);
            if (k == std::string::npos) throw std::runtime_error("Learner::LoadStats(): missing key " + key);
            k = j.find(':', k) + 1;
            while (k < j.size() && (j[k] == ' ' || j[k] == '[' || j[k] == '\n')) k++;
            return std::strtod(j.c_str() + k, nullptr);
        };
        totalTimesteps = (uint64_t)num("cumulative_timesteps");
        ppo->cumulativeModelUpdates = (uint64_t)num("cumulative_model_updates");
        totalEpochs = (uint64_t)num("epoch");
        const size_t rs = j.find("\"reward_running_stats\"");
        returnStats = WelfordRunningStat(1);
        returnStats.runningMean[0] = num("mean", rs); returnStats.runningVariance[0] = num("var", rs); returnStats.count = (int64_t)num("count", rs);
    }

private:
    static void MakeDirs(const std::string& path) {
        for (size_t i = 1; i <= path.size(); i++)
            if (i == path.size() || path[i] == '/') mkdir(path.substr(0, i).c_str(), 0777);
    }
    static std::vector<uint64_t> NumberedFolders(const std::string& folder) {
        std::vector<uint64_t> out;
        if (DIR* d = opendir(folder.c_str())) {
            while (dirent* ent = readdir(d)) {
                const std::string n = ent->d_name;
                if (!n.empty() && n.find_first_not_of("0123456789") == std::string::npos) out.push_back(std::strtoull(n.c_str(), nullptr, 10));
            }
            closedir(d);
        }
        std::sort(out.begin(), out.end());
        return out;
    }
    static void RemoveTree(const std::string& folder) {
        if (DIR* d = opendir(folder.c_str())) {
            while (dirent* ent = readdir(d)) {
                const std::string n = ent->d_name;
                if (n != "." && n != "..") remove((folder + "/" + n).c_str());
            }
            closedir(d);
        }
        rmdir(folder.c_str());
    }
};
}  // namespace RLGPC
