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
//   RLGPC::  PPOLearnerConfig, LearnerConfig (field for field), EnvCreateResult / EnvCreateFn (GameInst.h:10-14),
//            GameTrajectory (device views), ThreadAgentManager (ThreadAgentManager.h:10-60 surface:
//            CreateAgents / StartAgents / StopAgents / CollectTimesteps / GetMetrics / ResetMetrics)
//   RocketSim::Init(path)  — loads soccar/*.cmf exactly like R/RocketSim.cpp:70-212 and keeps the bytes for the engines
//
// Built-in plugins are recognised by dynamic_cast and become configuration of the fused device kernels
// (RLGB200::CfgFromMatch); their virtuals are never called on the host.  A user-defined StateSetter is supported
// through the host path (Arena/Car/Ball proxies -> rlg_engine_set_state + rlg_engine_reset_current).  User-defined
// OBSBuilder / RewardFunction / TerminalCondition / ActionParser subclasses and StepCallback need per-step GameState
// materialisation on the host; CfgFromMatch rejects them with std::runtime_error (RG_ERR_CLOSE style) — round-2 work.
// Errors: every failing C-ABI call is re-thrown as std::runtime_error(rlg_last_error()) like RG_ERR_CLOSE
// (G/Framework.h:17-22).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
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
    float Dot(const Vec& o) const { return x * o.x + y * o.y + z * o.z + _w * o._w; }
    float LengthSq() const { return Dot(*this); }
    float Length() const { float l = LengthSq(); return l > 0 ? std::sqrt(l) : 0; }
    Vec Normalized() const { float l = Length(); return l > 1e-12f ? Vec(x / l, y / l, z / l) : Vec(); }
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

// R/Sim/Car/Car.h:17-115 (the members a StateSetter normally writes; everything else keeps its default)
struct CarState {
    Vec pos{0, 0, 17.f};
    RotMat rotMat;
    Vec vel, angVel;
    bool isOnGround = true;
    bool hasJumped = false, hasDoubleJumped = false, hasFlipped = false;
    float boost = 100.f / 3;
    bool isDemoed = false;
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

struct Action {  // G/Utils/BasicTypes/Action.h:5-47
    float throttle = 0, steer = 0, pitch = 0, yaw = 0, roll = 0, jump = 0, boost = 0, handbrake = 0;
};
typedef std::vector<Action> ActionSet;

struct PhysObj {  // G/Utils/Gamestates/PhysObj.h
    Vec pos, vel, angVel;
    RotMat rotMat;
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
    bool hasFlip = false, ballTouchedStep = false, ballTouchedTick = false;
    float boostFraction = 0;
};
struct GameState {  // G/Utils/Gamestates/GameState.h:19-57
    int scoreLine[2] = {0, 0};
    int lastTouchCarID = -1;
    std::vector<PlayerData> players;
    PhysObj ball, ballInv;
    bool boostPads[RLG_NUM_PADS] = {}, boostPadsInv[RLG_NUM_PADS] = {};
    uint64_t lastTickCount = 0;
    float deltaTime = 0;
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

// ---- built-ins: configuration carriers for the fused device kernels -------------------------------------------------
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
class EventReward : public RewardFunction {  // CommonRewards.h:6-49
public:
    struct WeightScales {
        float goal = 0, teamGoal = 0, concede = 0, assist = 0, touch = 0, shot = 0, shotPass = 0, save = 0, demo = 0, demoed = 0, boostPickup = 0;
    };
    WeightScales weights;
    explicit EventReward(WeightScales w) : weights(w) {}
};
class VelocityReward : public RewardFunction {
public:
    bool isNegative;
    explicit VelocityReward(bool isNegative_ = false) : isNegative(isNegative_) {}
};
class SaveBoostReward : public RewardFunction {  // CommonRewards.h:61-70
public:
    float exponent;
    explicit SaveBoostReward(float exponent_ = 0.5f) : exponent(exponent_) {}
};
class TouchBallReward : public RewardFunction {  // CommonRewards.h:110-124
public:
    float aerialWeight;
    explicit TouchBallReward(float aerialWeight_ = 0) : aerialWeight(aerialWeight_) {}
};
class VelocityPlayerToBallReward : public RewardFunction {};
class FaceBallReward : public RewardFunction {};
class VelocityBallToGoalReward : public RewardFunction {
public:
    bool ownGoal;
    explicit VelocityBallToGoalReward(bool ownGoal_ = false) : ownGoal(ownGoal_) {}
};
class CombinedReward : public RewardFunction {  // CombinedReward.h
public:
    std::vector<RewardFunction*> rewardFuncs;
    std::vector<float> rewardWeights;
    bool ownsFuncs;
    CombinedReward(std::vector<std::pair<RewardFunction*, float>> funcsWithWeights, bool ownsFuncs_ = false) : ownsFuncs(ownsFuncs_) {
        for (auto& p : funcsWithWeights) { rewardFuncs.push_back(p.first); rewardWeights.push_back(p.second); }
    }
    ~CombinedReward() override { if (ownsFuncs) for (auto* f : rewardFuncs) delete f; }
};
class ZeroSumReward : public RewardFunction {  // ZeroSumReward.h:8-26
public:
    RewardFunction* childFunc;
    bool ownsFunc;
    float teamSpirit, opponentScale;
    ZeroSumReward(RewardFunction* child, float teamSpirit_, float opponentScale_ = 1, bool ownsFunc_ = false)
        : childFunc(child), ownsFunc(ownsFunc_), teamSpirit(teamSpirit_), opponentScale(opponentScale_) {}
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
    int64_t maxSteps;
    explicit NoTouchCondition(int64_t maxSteps_) : maxSteps(maxSteps_) {}
    bool IsTerminal(const GameState&) override { RLGB200_DEVICE_ONLY("NoTouchCondition"); }
};
class GoalScoreCondition : public TerminalCondition {
public:
    bool IsTerminal(const GameState&) override { RLGB200_DEVICE_ONLY("GoalScoreCondition"); }
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
    Match(RewardFunction* rewardFn_, std::vector<TerminalCondition*> terminalConditions_, OBSBuilder* obsBuilder_, ActionParser* actionParser_,
          StateSetter* stateSetter_, int teamSize_ = 1, bool spawnOpponents_ = true)
        : rewardFn(rewardFn_), terminalConditions(std::move(terminalConditions_)), obsBuilder(obsBuilder_), actionParser(actionParser_),
          stateSetter(stateSetter_), teamSize(teamSize_), spawnOpponents(spawnOpponents_), playerAmount(teamSize_ * (spawnOpponents_ ? 2 : 1)) {}
};
class Gym {  // G/Gym.h:8-31 — with the device engine a Gym is a description; stepping happens batched (RLGB200::BatchedGym)
public:
    Match* match;
    int tickSkip;
    struct StepResult {
        FList2 obs;
        FList reward;
        bool done;
    };
    RocketSim::CarConfig carConfig;
    RocketSim::MutatorConfig mutatorConfig;
    Gym(Match* match_, int tickSkip_, RocketSim::CarConfig carConfig_ = RocketSim::CAR_CONFIG_OCTANE,
        RocketSim::GameMode gameMode = RocketSim::GameMode::SOCCAR, RocketSim::MutatorConfig mutatorConfig_ = RocketSim::MutatorConfig(RocketSim::GameMode::SOCCAR))
        : match(match_), tickSkip(tickSkip_), carConfig(carConfig_), mutatorConfig(mutatorConfig_) {}  // G/Gym.h:18 (soccar only)
};
}  // namespace RLGSC

// ---------------------------------------------------------------------------------------------------------------------
namespace RLGB200 {
using namespace RLGSC;

// Built-in plugin objects -> rlg_engine_cfg ("recognises built-in plugin classes by dynamic_cast").
inline rlg_engine_cfg CfgFromMatch(const Match& m, int tickSkip, int numArenas, int device = 0, uint64_t seed = 123) {
    rlg_engine_cfg c;
    rlg_engine_cfg_default(&c);
    c.num_arenas = numArenas; c.team_size = m.teamSize; c.spawn_opponents = m.spawnOpponents ? 1 : 0; c.tick_skip = tickSkip;
    c.device = device; c.seed = seed;
    // obs
    if (dynamic_cast<DefaultOBS*>(m.obsBuilder)) c.obs_kind = RLG_OBS_DEFAULT;
    else if (auto* p = dynamic_cast<DefaultOBSPadded*>(m.obsBuilder)) { c.obs_kind = RLG_OBS_PADDED; c.obs_max_players = p->maxPlayers; }
    else throw std::runtime_error("RLGB200: user-defined OBSBuilder needs the host-plugin path (not built yet); use DefaultOBS / DefaultOBSPadded");
    // action parser
    if (!dynamic_cast<DiscreteAction*>(m.actionParser)) throw std::runtime_error("RLGB200: only DiscreteAction is fused on the device");
    // rewards
    RewardFunction* rf = m.rewardFn;
    c.zero_sum = 0;
    if (auto* z = dynamic_cast<ZeroSumReward*>(rf)) { c.zero_sum = 1; c.team_spirit = z->teamSpirit; c.opponent_scale = z->opponentScale; rf = z->childFunc; }
    std::vector<std::pair<RewardFunction*, float>> terms;
    if (auto* cr = dynamic_cast<CombinedReward*>(rf)) for (size_t i = 0; i < cr->rewardFuncs.size(); i++) terms.push_back({cr->rewardFuncs[i], cr->rewardWeights[i]});
    else terms.push_back({rf, 1.f});
    if (terms.size() > RLG_MAX_REWARD_TERMS) throw std::runtime_error("RLGB200: more than RLG_MAX_REWARD_TERMS reward terms");
    c.num_reward_terms = (int32_t)terms.size();
    for (size_t i = 0; i < terms.size(); i++) {
        rlg_reward_term& t = c.reward_terms[i];
        memset(&t, 0, sizeof(t));
        t.weight = terms[i].second;
        RewardFunction* f = terms[i].first;
        if (auto* e = dynamic_cast<EventReward*>(f)) {
            t.kind = RLG_REW_EVENT;
            const auto& w = e->weights;
            const float v[11] = {w.goal, w.teamGoal, w.concede, w.assist, w.touch, w.shot, w.shotPass, w.save, w.demo, w.demoed, w.boostPickup};
            for (int k = 0; k < 11; k++) t.params[k] = v[k];
        } else if (dynamic_cast<VelocityPlayerToBallReward*>(f)) t.kind = RLG_REW_VEL_PLAYER_TO_BALL;
        else if (auto* g = dynamic_cast<VelocityBallToGoalReward*>(f)) { t.kind = RLG_REW_VEL_BALL_TO_GOAL; t.params[0] = g->ownGoal ? 1.f : 0.f; }
        else if (dynamic_cast<FaceBallReward*>(f)) t.kind = RLG_REW_FACE_BALL;
        else if (auto* v = dynamic_cast<VelocityReward*>(f)) { t.kind = RLG_REW_VELOCITY; t.params[0] = v->isNegative ? 1.f : 0.f; }
        else if (auto* sb = dynamic_cast<SaveBoostReward*>(f)) { t.kind = RLG_REW_SAVE_BOOST; t.params[0] = sb->exponent; }
        else if (auto* tb = dynamic_cast<TouchBallReward*>(f)) { t.kind = RLG_REW_TOUCH_BALL; t.params[0] = tb->aerialWeight; }
        else throw std::runtime_error("RLGB200: user-defined RewardFunction needs the host-plugin path (not built yet)");
    }
    // terminals
    c.no_touch_max_steps = 0; c.goal_score_terminal = 0;
    for (auto* tc : m.terminalConditions) {
        if (auto* nt = dynamic_cast<NoTouchCondition*>(tc)) c.no_touch_max_steps = (int32_t)nt->maxSteps;
        else if (dynamic_cast<GoalScoreCondition*>(tc)) c.goal_score_terminal = 1;
        else throw std::runtime_error("RLGB200: user-defined TerminalCondition needs the host-plugin path (not built yet)");
    }
    // state setter
    if (dynamic_cast<KickoffState*>(m.stateSetter)) c.state_setter = RLG_SETTER_KICKOFF;
    else if (auto* rs = dynamic_cast<RandomState*>(m.stateSetter)) {
        c.state_setter = RLG_SETTER_RANDOM; c.rand_ball_speed = rs->randBallSpeed; c.rand_car_speed = rs->randCarSpeed; c.cars_on_ground = rs->carsOnGround;
    } else c.state_setter = RLG_SETTER_HOST;  // user StateSetter::ResetState(Arena*) runs on the host
    return c;
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

// Runs a user StateSetter on the host for the arenas in `ids` and uploads the result (StateSetter::ResetState(Arena*),
// G/Envs/Match.cpp:54-70): cars/ball the setter did not touch keep the default kickoff-spawn state.
inline void RunHostStateSetter(Engine& e, StateSetter& setter, const std::vector<int32_t>& ids, float* obsOut = nullptr) {
    const int P = e.NumPlayers(), n = (int)ids.size();
    if (n == 0) return;
    std::vector<rlg_car_state> cars((size_t)n * P);
    std::vector<rlg_ball_state> balls(n);
    std::vector<rlg_pad_state> pads((size_t)n * RLG_NUM_PADS);
    std::vector<int64_t> ticks(n, -1);
    Check(rlg_engine_get_state(e.h, ids.data(), n, cars.data(), balls.data(), pads.data(), nullptr));
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
        setter.ResetState(&arena);
        for (int c = 0; c < P; c++) {
            rlg_car_state& o = cars[(size_t)i * P + c];
            const CarState& s = carObjs[c].state;
            int32_t id = o.car_id, team = o.team;
            memset(&o, 0, sizeof(o));
            o.car_id = id ? id : c + 1; o.team = team;
            o.pos[0] = s.pos.x; o.pos[1] = s.pos.y; o.pos[2] = s.pos.z;
            const Vec* cols[3] = {&s.rotMat.forward, &s.rotMat.right, &s.rotMat.up};
            float* dst[3] = {o.rot_forward, o.rot_right, o.rot_up};
            for (int k = 0; k < 3; k++) { dst[k][0] = cols[k]->x; dst[k][1] = cols[k]->y; dst[k][2] = cols[k]->z; }
            o.vel[0] = s.vel.x; o.vel[1] = s.vel.y; o.vel[2] = s.vel.z;
            o.ang_vel[0] = s.angVel.x; o.ang_vel[1] = s.angVel.y; o.ang_vel[2] = s.angVel.z;
            o.is_on_ground = s.isOnGround; o.has_jumped = s.hasJumped; o.has_double_jumped = s.hasDoubleJumped; o.has_flipped = s.hasFlipped;
            o.boost = s.boost; o.is_demoed = s.isDemoed;
            o.hit_tick = -1; o.hit_extra_tick = -1;
        }
        balls[i].pos[0] = ball.state.pos.x; balls[i].pos[1] = ball.state.pos.y; balls[i].pos[2] = ball.state.pos.z;
        balls[i].vel[0] = ball.state.vel.x; balls[i].vel[1] = ball.state.vel.y; balls[i].vel[2] = ball.state.vel.z;
        balls[i].ang_vel[0] = ball.state.angVel.x; balls[i].ang_vel[1] = ball.state.angVel.y; balls[i].ang_vel[2] = ball.state.angVel.z;
        for (int p = 0; p < RLG_NUM_PADS; p++) { auto& ps = pads[(size_t)i * RLG_NUM_PADS + p]; ps.is_active = 1; ps.cooldown = 0; ps.prev_locked_car_id = 0; }  // Match.cpp:66-67
    }
    Check(rlg_engine_set_state(e.h, ids.data(), n, cars.data(), balls.data(), pads.data(), ticks.data()));
    std::vector<uint8_t> mask(e.NumArenas(), 0);
    for (int32_t id : ids) mask[id] = 1;
    if (obsOut) Check(rlg_engine_reset_current_to(e.h, mask.data(), obsOut, nullptr));
    else Check(rlg_engine_reset_current(e.h, mask.data(), nullptr));
    Check(rlg_engine_sync(e.h));
}
}  // namespace RLGB200

// ---------------------------------------------------------------------------------------------------------------------
namespace RLGPC {
using RLGSC::IList;
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
};
struct EnvCreateResult {  // GameInst.h:10-13
    RLGSC::Match* match;
    RLGSC::Gym* gym;
};
typedef std::function<EnvCreateResult()> EnvCreateFn;
typedef std::map<std::string, double> Report;

// GameTrajectory.h:5-18 as DEVICE views in the reference's concatenated row order (size rows)
struct GameTrajectory {
    uint64_t size = 0;
    int obsSize = 0;
    float* states = nullptr; int64_t* actions = nullptr; float* logProbs = nullptr; float* rewards = nullptr; float* nextStates = nullptr;
    float* dones = nullptr; float* truncateds = nullptr;
    float* valueTargets = nullptr; float* advantages = nullptr;  // filled when gae = true
};

// ThreadAgentManager.h:10-60 over ONE device engine: amount x gamesPerAgent arenas, no host threads.
class ThreadAgentManager {
public:
    LearnerConfig cfg;
    int device;
    std::unique_ptr<RLGB200::Engine> engine;
    rlg_collector* collector = nullptr;
    EnvCreateResult probe{nullptr, nullptr};
    int stepsPerCollect = 0;
    bool disableCollection = false;
    double lastIterationTime = 0;
    ThreadAgentManager(const LearnerConfig& cfg_, int device_ = 0) : cfg(cfg_), device(device_) {}
    ~ThreadAgentManager() {
        if (collector) rlg_collector_destroy(collector);
        if (probe.gym) delete probe.gym;
        if (probe.match) delete probe.match;
    }
    void CreateAgents(EnvCreateFn func, int amount, int gamesPerAgent) {
        probe = func();  // one Match/Gym to read the plugin configuration from (ThreadAgent.cpp:197-206 makes one per game)
        rlg_engine_cfg ec = RLGB200::CfgFromMatch(*probe.match, probe.gym->tickSkip, amount * gamesPerAgent, device, (uint64_t)cfg.randomSeed);
        ec.car_preset = probe.gym->carConfig.preset;
        ec.mutators = probe.gym->mutatorConfig.ToC();  // Gym.cpp:43 arena->SetMutatorConfig
        ec.mutators_set = 1;
        engine.reset(new RLGB200::Engine(ec));
        engine->LoadMeshes(RocketSim::CollisionMeshBlobs());
        const int N = engine->NumArenas() * engine->NumPlayers();
        stepsPerCollect = (int)std::max<int64_t>(1, (cfg.timestepsPerIteration + N - 1) / N);
        rlg_collector_cfg cc;
        memset(&cc, 0, sizeof(cc));
        if (cfg.ppo.policyLayerSizes.size() != cfg.ppo.criticLayerSizes.size() || cfg.ppo.policyLayerSizes.size() > RLG_MAX_HIDDEN_LAYERS)
            throw std::runtime_error("RLGB200: policy/critic need the same number (<= 4) of hidden layers");
        cc.num_hidden = (int32_t)cfg.ppo.policyLayerSizes.size();
        for (int i = 0; i < cc.num_hidden; i++) { cc.policy_hidden[i] = cfg.ppo.policyLayerSizes[i]; cc.critic_hidden[i] = cfg.ppo.criticLayerSizes[i]; }
        cc.max_steps = stepsPerCollect; cc.seed = (uint64_t)cfg.randomSeed; cc.temperature = cfg.ppo.policyTemperature; cc.deterministic = cfg.deterministic;
        RLGB200::Check(rlg_collector_create(engine->h, &cc, &collector));
        if (ec.state_setter == RLG_SETTER_HOST) RLGB200::Check(rlg_collector_set_reset_hook(collector, &ThreadAgentManager::ResetHook, this));
    }
    static void ResetHook(void* user, const int32_t* ids, int n, float* obsOut) {
        auto* self = static_cast<ThreadAgentManager*>(user);
        RLGB200::RunHostStateSetter(*self->engine, *self->probe.match->stateSetter, std::vector<int32_t>(ids, ids + n), obsOut);
    }
    // torch nn.Linear tensors of DiscretePolicy (net 0) / ValueEstimator (net 1): PPOLearner pushes them after every update
    void SetLayer(int net, int layer, const float* W, const float* b, int outDim, int inDim) {
        RLGB200::Check(rlg_collector_set_layer(collector, net, layer, W, b, outDim, inDim));
    }
    void StartAgents() {
        if (engine->cfg.state_setter == RLG_SETTER_HOST) {
            std::vector<int32_t> ids(engine->NumArenas());
            for (size_t i = 0; i < ids.size(); i++) ids[i] = (int32_t)i;
            RLGB200::RunHostStateSetter(*engine, *probe.match->stateSetter, ids);
        } else {
            RLGB200::Check(rlg_engine_reset(engine->h, nullptr, nullptr));
        }
    }
    void StopAgents() { RLGB200::Check(rlg_engine_sync(engine->h)); }
    // Blocks until >= amount player-steps are collected (ThreadAgentManager.cpp:16-80); returns T-major device views.
    rlg_traj_view CollectTimesteps(uint64_t amount, bool computeGae = true, float returnStd = 1.f) {
        const uint64_t N = (uint64_t)engine->NumArenas() * engine->NumPlayers();
        int steps = (int)std::max<uint64_t>(1, (amount + N - 1) / N);
        if (steps > stepsPerCollect) throw std::runtime_error("CollectTimesteps: amount exceeds timestepsPerIteration");
        RLGB200::Check(rlg_collector_collect(collector, steps, nullptr));
        if (computeGae) RLGB200::Check(rlg_collector_gae(collector, cfg.gaeGamma, cfg.gaeLambda, returnStd, cfg.rewardClipRange, nullptr));
        RLGB200::Check(rlg_engine_sync(engine->h));
        rlg_traj_view v;
        RLGB200::Check(rlg_collector_view(collector, &v));
        return v;
    }
    void GetMetrics(Report& report) {
        double sm = 0, im = 0; int32_t sn = 0, in = 0;
        RLGB200::Check(rlg_collector_kernel_times(collector, &sm, &sn, &im, &in));
        rlg_metrics_host m;
        RLGB200::Check(rlg_engine_metrics(engine->h, &m));
        report["Average Step Reward"] = m.avg_step_reward;
        report["Average Episode Reward"] = m.avg_episode_reward;
        report["Env Step Time"] = sm * 1e-3;
        report["Policy Infer Time"] = im * 1e-3;
    }
    void ResetMetrics() { RLGB200::Check(rlg_engine_reset_metrics(engine->h)); }
};
}  // namespace RLGPC
