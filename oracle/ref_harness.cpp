// ref_harness.cpp — TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// A thin C-ABI harness around the UNMODIFIED reference (RocketSim + RLGymSim_CPP
// compiled from /root/reference by oracle/Makefile into oracle/_ref/librlref.so).
// It exposes exactly what the parity tests and bench.py's cpu_baseline /
// `--impl reference` leg need:
//   * raw Arena access: inject CarState/BallState/pad state, Arena::Step(n), read back
//   * Gym access: Gym::Reset / Gym::Step with the built-in plugins configured from
//     the same rlg_engine_cfg the CUDA engine takes
//   * a multithreaded Gym::Step loop that mirrors GameInst::Step (the CPU baseline)
// State structs are the ones in include/rlgym_b200.h so both sides speak one format.
//
// Reference entry points used (all public API of the reference):
//   RocketSim::InitFromMem            RocketSim/src/RocketSim.cpp:105
//   Arena::Create/AddCar/Step         RocketSim/src/Sim/Arena/Arena.cpp:579,47,716
//   Car::SetState/GetState            RocketSim/src/Sim/Car/Car.cpp:23,10
//   Gym::Reset/Step                   src/RLGymSim_CPP/Gym.cpp:58,68
#include <RLGymSim_CPP/Gym.h>
#include <RLGymSim_CPP/Utils/OBSBuilders/DefaultOBS.h>
#include <RLGymSim_CPP/Utils/OBSBuilders/DefaultOBSPadded.h>
#include <RLGymSim_CPP/Utils/RewardFunctions/CombinedReward.h>
#include <RLGymSim_CPP/Utils/RewardFunctions/CommonRewards.h>
#include <RLGymSim_CPP/Utils/RewardFunctions/ZeroSumReward.h>
#include <RLGymSim_CPP/Utils/ActionParsers/DiscreteAction.h>
#include <RLGymSim_CPP/Utils/StateSetters/KickoffState.h>
#include <RLGymSim_CPP/Utils/StateSetters/RandomState.h>
#include <RLGymSim_CPP/Utils/TerminalConditions/NoTouchCondition.h>
#include <RLGymSim_CPP/Utils/TerminalConditions/GoalScoreCondition.h>

#include <RocketSim/libsrc/bullet3-3.24/BulletCollision/CollisionShapes/btBoxShape.h>
#include <RocketSim/libsrc/bullet3-3.24/BulletCollision/CollisionShapes/btSphereShape.h>
#include <RocketSim/libsrc/bullet3-3.24/BulletCollision/CollisionShapes/btTriangleShape.h>
#include <RocketSim/libsrc/bullet3-3.24/BulletCollision/CollisionDispatch/btBoxBoxDetector.h>
#include <RocketSim/libsrc/bullet3-3.24/BulletCollision/NarrowPhaseCollision/btGjkEpaPenetrationDepthSolver.h>
#include <RocketSim/libsrc/bullet3-3.24/BulletCollision/NarrowPhaseCollision/btGjkPairDetector.h>
#include <RocketSim/libsrc/bullet3-3.24/BulletCollision/NarrowPhaseCollision/btVoronoiSimplexSolver.h>

#include "../include/rlgym_b200.h"

#include <atomic>
#include <chrono>
#include <thread>

using namespace RLGSC;

namespace {

static void copy3(float* dst, const Vec& v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }
static Vec vec3(const float* s) { return Vec(s[0], s[1], s[2]); }

static Car* car_by_id(Arena* a, uint32_t id) {
    auto it = a->_carIDMap.find(id);
    return it == a->_carIDMap.end() ? nullptr : it->second;
}

static void controls_to_pod(const CarControls& c, rlg_controls* o) {
    o->throttle = c.throttle; o->steer = c.steer; o->pitch = c.pitch; o->yaw = c.yaw; o->roll = c.roll;
    o->jump = c.jump; o->boost = c.boost; o->handbrake = c.handbrake;
}
static CarControls controls_from_pod(const rlg_controls& c) {
    CarControls o;
    o.throttle = c.throttle; o.steer = c.steer; o.pitch = c.pitch; o.yaw = c.yaw; o.roll = c.roll;
    o.jump = c.jump != 0; o.boost = c.boost != 0; o.handbrake = c.handbrake != 0;
    return o;
}

static void car_to_pod(Car* car, rlg_car_state* o) {
    CarState s = car->GetState();
    memset(o, 0, sizeof(*o));
    copy3(o->pos, s.pos);
    copy3(o->rot_forward, s.rotMat.forward); copy3(o->rot_right, s.rotMat.right); copy3(o->rot_up, s.rotMat.up);
    copy3(o->vel, s.vel); copy3(o->ang_vel, s.angVel);
    o->is_on_ground = s.isOnGround;
    for (int i = 0; i < 4; i++) o->wheels_with_contact[i] = s.wheelsWithContact[i];
    o->has_jumped = s.hasJumped; o->has_double_jumped = s.hasDoubleJumped; o->has_flipped = s.hasFlipped;
    copy3(o->flip_rel_torque, s.flipRelTorque);
    o->jump_time = s.jumpTime; o->flip_time = s.flipTime;
    o->is_flipping = s.isFlipping; o->is_jumping = s.isJumping;
    o->air_time = s.airTime; o->air_time_since_jump = s.airTimeSinceJump;
    o->boost = s.boost; o->time_spent_boosting = s.timeSpentBoosting;
    o->is_supersonic = s.isSupersonic; o->supersonic_time = s.supersonicTime; o->handbrake_val = s.handbrakeVal;
    o->is_auto_flipping = s.isAutoFlipping; o->auto_flip_timer = s.autoFlipTimer; o->auto_flip_torque_scale = s.autoFlipTorqueScale;
    o->world_contact_has = s.worldContact.hasContact; copy3(o->world_contact_normal, s.worldContact.contactNormal);
    o->car_contact_other_id = (int32_t)s.carContact.otherCarID; o->car_contact_cooldown = s.carContact.cooldownTimer;
    o->is_demoed = s.isDemoed; o->demo_respawn_timer = s.demoRespawnTimer;
    o->hit_valid = s.ballHitInfo.isValid;
    copy3(o->hit_rel_pos_on_ball, s.ballHitInfo.relativePosOnBall);
    copy3(o->hit_ball_pos, s.ballHitInfo.ballPos);
    copy3(o->hit_extra_vel, s.ballHitInfo.extraHitVel);
    o->hit_tick = (int64_t)s.ballHitInfo.tickCountWhenHit;
    o->hit_extra_tick = (int64_t)s.ballHitInfo.tickCountWhenExtraImpulseApplied;
    controls_to_pod(s.lastControls, &o->last_controls);
    auto& wi = car->_bulletVehicle.m_wheelInfo;
    o->wheel_steer_angle = wi[0].m_steerAngle;
    o->wheel_engine_force = wi[0].m_engineForce;
    o->wheel_brake = wi[0].m_brake;
    for (int i = 0; i < 4; i++) {
        o->wheel_lat_friction[i] = wi[i].m_latFriction;
        o->wheel_long_friction[i] = wi[i].m_longFriction;
        o->wheel_extra_pushback[i] = wi[i].m_extraPushback;
    }
    o->car_id = (int32_t)car->id;
    o->team = (int32_t)car->team;
}

static void car_from_pod(Car* car, const rlg_car_state& i) {
    CarState s;
    s.pos = vec3(i.pos);
    s.rotMat = RotMat(vec3(i.rot_forward), vec3(i.rot_right), vec3(i.rot_up));
    s.vel = vec3(i.vel); s.angVel = vec3(i.ang_vel);
    s.isOnGround = i.is_on_ground;
    for (int k = 0; k < 4; k++) s.wheelsWithContact[k] = i.wheels_with_contact[k];
    s.hasJumped = i.has_jumped; s.hasDoubleJumped = i.has_double_jumped; s.hasFlipped = i.has_flipped;
    s.flipRelTorque = vec3(i.flip_rel_torque);
    s.jumpTime = i.jump_time; s.flipTime = i.flip_time;
    s.isFlipping = i.is_flipping; s.isJumping = i.is_jumping;
    s.airTime = i.air_time; s.airTimeSinceJump = i.air_time_since_jump;
    s.boost = i.boost; s.timeSpentBoosting = i.time_spent_boosting;
    s.isSupersonic = i.is_supersonic; s.supersonicTime = i.supersonic_time; s.handbrakeVal = i.handbrake_val;
    s.isAutoFlipping = i.is_auto_flipping; s.autoFlipTimer = i.auto_flip_timer; s.autoFlipTorqueScale = i.auto_flip_torque_scale;
    s.worldContact.hasContact = i.world_contact_has; s.worldContact.contactNormal = vec3(i.world_contact_normal);
    s.carContact.otherCarID = (uint32_t)i.car_contact_other_id; s.carContact.cooldownTimer = i.car_contact_cooldown;
    s.isDemoed = i.is_demoed; s.demoRespawnTimer = i.demo_respawn_timer;
    s.ballHitInfo.isValid = i.hit_valid;
    s.ballHitInfo.relativePosOnBall = vec3(i.hit_rel_pos_on_ball);
    s.ballHitInfo.ballPos = vec3(i.hit_ball_pos);
    s.ballHitInfo.extraHitVel = vec3(i.hit_extra_vel);
    s.ballHitInfo.tickCountWhenHit = (uint64_t)i.hit_tick;
    s.ballHitInfo.tickCountWhenExtraImpulseApplied = (uint64_t)i.hit_extra_tick;
    s.lastControls = controls_from_pod(i.last_controls);
    car->SetState(s);
    auto& wi = car->_bulletVehicle.m_wheelInfo;
    wi[0].m_steerAngle = wi[1].m_steerAngle = i.wheel_steer_angle;
    for (int k = 0; k < 4; k++) {
        wi[k].m_engineForce = i.wheel_engine_force;
        wi[k].m_brake = i.wheel_brake;
        wi[k].m_latFriction = i.wheel_lat_friction[k];
        wi[k].m_longFriction = i.wheel_long_friction[k];
        wi[k].m_extraPushback = i.wheel_extra_pushback[k];
    }
}

struct RefGym {
    Gym* gym = nullptr;
    Match* match = nullptr;
    RewardFunction* reward = nullptr;
    std::vector<TerminalCondition*> terms;
    OBSBuilder* obs = nullptr;
    ActionParser* parser = nullptr;
    StateSetter* setter = nullptr;
    ~RefGym() {
        delete gym; delete match; delete reward;
        for (auto t : terms) delete t;
        delete obs; delete parser; delete setter;
    }
};

static MutatorConfig mutators_from(const rlg_engine_cfg* cfg) {
    MutatorConfig m(GameMode::SOCCAR);
    if (!cfg || !cfg->mutators_set) return m;
    const rlg_mutators& u = cfg->mutators;
    m.gravity = Vec(u.gravity[0], u.gravity[1], u.gravity[2]);
    m.carMass = u.car_mass; m.carWorldFriction = u.car_world_friction; m.carWorldRestitution = u.car_world_restitution;
    m.ballMass = u.ball_mass; m.ballMaxSpeed = u.ball_max_speed; m.ballDrag = u.ball_drag;
    m.ballWorldFriction = u.ball_world_friction; m.ballWorldRestitution = u.ball_world_restitution;
    m.jumpAccel = u.jump_accel; m.jumpImmediateForce = u.jump_immediate_force;
    m.boostAccelGround = u.boost_accel_ground; m.boostAccelAir = u.boost_accel_air; m.boostUsedPerSecond = u.boost_used_per_second;
    m.respawnDelay = u.respawn_delay; m.bumpCooldownTime = u.bump_cooldown_time;
    m.boostPadCooldown_Big = u.boost_pad_cooldown_big; m.boostPadCooldown_Small = u.boost_pad_cooldown_small;
    m.carSpawnBoostAmount = u.car_spawn_boost_amount; m.ballHitExtraForceScale = u.ball_hit_extra_force_scale; m.bumpForceScale = u.bump_force_scale;
    m.ballRadius = u.ball_radius; m.unlimitedFlips = u.unlimited_flips != 0; m.unlimitedDoubleJumps = u.unlimited_double_jumps != 0;
    m.demoMode = (DemoMode)u.demo_mode; m.enableTeamDemos = u.enable_team_demos != 0; m.goalBaseThresholdY = u.goal_base_threshold_y;
    return m;
}
static const CarConfig& preset_config(int preset) {
    switch (preset) {
    case RLG_CAR_DOMINUS: return CAR_CONFIG_DOMINUS;
    case RLG_CAR_PLANK: return CAR_CONFIG_PLANK;
    case RLG_CAR_BREAKOUT: return CAR_CONFIG_BREAKOUT;
    case RLG_CAR_HYBRID: return CAR_CONFIG_HYBRID;
    case RLG_CAR_MERC: return CAR_CONFIG_MERC;
    }
    return CAR_CONFIG_OCTANE;
}
static RewardFunction* make_term(const rlg_reward_term& t) {
    switch (t.kind) {
    case RLG_REW_EVENT: {
        EventReward::WeightScales w;
        for (int i = 0; i < 11; i++) w[i] = t.params[i];
        return new EventReward(w);
    }
    case RLG_REW_VEL_PLAYER_TO_BALL: return new VelocityPlayerToBallReward();
    case RLG_REW_VEL_BALL_TO_GOAL: return new VelocityBallToGoalReward(t.params[0] != 0);
    case RLG_REW_FACE_BALL: return new FaceBallReward();
    case RLG_REW_VELOCITY: return new VelocityReward(t.params[0] != 0);
    case RLG_REW_SAVE_BOOST: return new SaveBoostReward(t.params[0]);
    case RLG_REW_TOUCH_BALL: return new TouchBallReward(t.params[0]);
    }
    return nullptr;
}

static RefGym* make_gym(const rlg_engine_cfg* cfg) {
    RefGym* g = new RefGym();
    std::vector<RewardFunction*> fns; std::vector<float> ws;
    for (int i = 0; i < cfg->num_reward_terms; i++) {
        fns.push_back(make_term(cfg->reward_terms[i]));
        ws.push_back(cfg->reward_terms[i].weight);
    }
    RewardFunction* combined = new CombinedReward(fns, ws, true);
    g->reward = cfg->zero_sum ? (RewardFunction*)new ZeroSumReward(combined, cfg->team_spirit, cfg->opponent_scale, true) : combined;
    if (cfg->no_touch_max_steps > 0) g->terms.push_back(new NoTouchCondition(cfg->no_touch_max_steps));
    if (cfg->goal_score_terminal) g->terms.push_back(new GoalScoreCondition());
    g->obs = cfg->obs_kind == RLG_OBS_PADDED ? (OBSBuilder*)new DefaultOBSPadded(cfg->obs_max_players) : (OBSBuilder*)new DefaultOBS();
    g->parser = new DiscreteAction();
    g->setter = cfg->state_setter == RLG_SETTER_KICKOFF
        ? (StateSetter*)new KickoffState()
        : (StateSetter*)new RandomState(cfg->rand_ball_speed, cfg->rand_car_speed, cfg->cars_on_ground);
    g->match = new Match(g->reward, g->terms, g->obs, g->parser, g->setter, cfg->team_size, cfg->spawn_opponents != 0);
    g->gym = new Gym(g->match, cfg->tick_skip, preset_config(cfg->car_preset), GameMode::SOCCAR, mutators_from(cfg));
    return g;
}

static int flatten_obs(const FList2& obs, float* out) {
    int k = 0;
    for (auto& row : obs) for (float v : row) out[k++] = v;
    return obs.empty() ? 0 : (int)obs[0].size();
}

} // namespace

extern "C" {

int ref_init(const void* const* blobs, const size_t* sizes, int n) {
    try {
        std::map<GameMode, std::vector<RocketSim::FileData>> m;
        auto& v = m[GameMode::SOCCAR];
        for (int i = 0; i < n; i++) {
            const byte* b = (const byte*)blobs[i];
            v.emplace_back(b, b + sizes[i]);
        }
        RocketSim::InitFromMem(m, true);
        return 0;
    } catch (std::exception&) { return -1; }
}

void ref_seed(uint32_t seed) { RocketSim::Math::GetRandEngine().seed(seed); }

// ---- raw arena -------------------------------------------------------------
// arena with the cfg's car preset and mutators (cars added in Gym::Gym order, mutators set like Gym.cpp:43)
void* ref_arena_create_cfg(const rlg_engine_cfg* cfg) {
    Arena* a = Arena::Create(GameMode::SOCCAR);
    a->SetMutatorConfig(mutators_from(cfg));
    for (int i = 0; i < cfg->team_size; i++) {
        a->AddCar(Team::BLUE, preset_config(cfg->car_preset));
        if (cfg->spawn_opponents) a->AddCar(Team::ORANGE, preset_config(cfg->car_preset));
    }
    return a;
}
void* ref_arena_create_preset(int team_size, int spawn_opponents, int preset) {
    Arena* a = Arena::Create(GameMode::SOCCAR);
    for (int i = 0; i < team_size; i++) {  // same order as Gym::Gym, Gym.cpp:46-50
        a->AddCar(Team::BLUE, preset_config(preset));
        if (spawn_opponents) a->AddCar(Team::ORANGE, preset_config(preset));
    }
    return a;
}
void* ref_arena_create(int team_size, int spawn_opponents) { return ref_arena_create_preset(team_size, spawn_opponents, 0); }
void ref_arena_destroy(void* h) { delete (Arena*)h; }
int ref_arena_num_cars(void* h) { return (int)((Arena*)h)->_cars.size(); }

// cars indexed by car id - 1
void ref_arena_set_state(void* h, const rlg_car_state* cars, const rlg_ball_state* ball,
                         const rlg_pad_state* pads, int64_t tick_count) {
    Arena* a = (Arena*)h;
    if (cars) {
        int n = (int)a->_cars.size();
        for (int i = 0; i < n; i++) car_from_pod(car_by_id(a, i + 1), cars[i]);
    }
    if (ball) {
        BallState b;
        b.pos = vec3(ball->pos); b.vel = vec3(ball->vel); b.angVel = vec3(ball->ang_vel);
        a->ball->SetState(b);
    }
    if (pads) {
        for (size_t i = 0; i < a->_boostPads.size(); i++) {
            BoostPadState s;
            s.isActive = pads[i].is_active; s.cooldown = pads[i].cooldown;
            s.prevLockedCarID = pads[i].prev_locked_car_id;
            a->_boostPads[i]->SetState(s);
        }
    }
    if (tick_count >= 0) a->tickCount = (uint64_t)tick_count;
}

void ref_arena_get_state(void* h, rlg_car_state* cars, rlg_ball_state* ball, rlg_pad_state* pads,
                         int64_t* tick_count) {
    Arena* a = (Arena*)h;
    if (cars) {
        int n = (int)a->_cars.size();
        for (int i = 0; i < n; i++) car_to_pod(car_by_id(a, i + 1), &cars[i]);
    }
    if (ball) {
        BallState b = a->ball->GetState();
        copy3(ball->pos, b.pos); copy3(ball->vel, b.vel); copy3(ball->ang_vel, b.angVel);
    }
    if (pads) {
        for (size_t i = 0; i < a->_boostPads.size(); i++) {
            BoostPadState s = a->_boostPads[i]->GetState();
            pads[i].is_active = s.isActive; pads[i].cooldown = s.cooldown;
            pads[i].prev_locked_car_id = (int32_t)s.prevLockedCarID;
        }
    }
    if (tick_count) *tick_count = (int64_t)a->tickCount;
}

void ref_arena_step(void* h, const rlg_controls* controls, int nticks) {
    Arena* a = (Arena*)h;
    if (controls) {
        int n = (int)a->_cars.size();
        for (int i = 0; i < n; i++) car_by_id(a, i + 1)->controls = controls_from_pod(controls[i]);
    }
    a->Step(nticks);
}

// debug: contact points of the last tick (manifolds live until the next tick's broadphase pass).
// out rows: [bodyA, bodyB, posA(3), posB(3), normal(3), dist, appliedImpulse, friction, restitution, lateralImpulse]
// body codes: -1 static, 0 ball, car id otherwise.
int ref_arena_dump_contacts(void* h, float* out, int maxRows) {
    Arena* a = (Arena*)h;
    auto code = [&](const btCollisionObject* o) -> float {
        if (o->getUserIndex() == BT_USERINFO_TYPE_BALL) return 0.f;
        if (o->getUserIndex() == BT_USERINFO_TYPE_CAR) return (float)((Car*)o->getUserPointer())->id;
        return -1.f;
    };
    int n = 0;
    auto* disp = a->_bulletWorld.getDispatcher();
    for (int i = 0; i < disp->getNumManifolds(); i++) {
        btPersistentManifold* m = disp->getManifoldByIndexInternal(i);
        for (int j = 0; j < m->getNumContacts() && n < maxRows; j++) {
            const btManifoldPoint& p = m->getContactPoint(j);
            float* r = out + n * 16;
            r[0] = code(m->getBody0()); r[1] = code(m->getBody1());
            for (int k = 0; k < 3; k++) { r[2 + k] = p.m_positionWorldOnA[k]; r[5 + k] = p.m_positionWorldOnB[k]; r[8 + k] = p.m_normalWorldOnB[k]; }
            r[11] = p.m_distance1; r[12] = p.m_appliedImpulse; r[13] = p.m_combinedFriction; r[14] = p.m_combinedRestitution; r[15] = p.m_appliedImpulseLateral1;
            n++;
        }
    }
    return n;
}



// wheel rays of the last tick: rows {contact point xyz, normal xyz, suspension length, in contact, ground is static} per wheel
void ref_arena_dump_wheels(void* h, int carIdx, float* out) {
    Car* car = car_by_id((Arena*)h, carIdx + 1);
    for (int i = 0; i < 4; i++) {
        const btWheelInfoRL& w = car->_bulletVehicle.m_wheelInfo[i];
        float* r = out + 9 * i;
        for (int k = 0; k < 3; k++) { r[k] = w.m_raycastInfo.m_contactPointWS[k]; r[3 + k] = w.m_raycastInfo.m_contactNormalWS[k]; }
        r[6] = w.m_raycastInfo.m_suspensionLength; r[7] = w.m_raycastInfo.m_isInContact ? 1.f : 0.f; r[8] = w.m_isInContactWithWorld ? -1.f : 0.f;
    }
}

// ---- contact-added trace: every point btManifoldResult::addContactPoint hands to the reference's callback, in call order ----
static ContactAddedCallback g_prev_added = nullptr;
static float g_trace[256][12];
static int g_trace_n = 0;
static bool trace_added(btManifoldPoint& cp, const btCollisionObjectWrapper* o0, int part0, int idx0, const btCollisionObjectWrapper* o1, int part1, int idx1) {
    if (g_trace_n < 256) {
        float* r = g_trace[g_trace_n++];
        auto code = [&](const btCollisionObject* o) -> float {
            if (o->getUserIndex() == BT_USERINFO_TYPE_BALL) return 0.f;
            if (o->getUserIndex() == BT_USERINFO_TYPE_CAR) return (float)((Car*)o->getUserPointer())->id;
            return -1.f;
        };
        r[0] = code(o0->m_collisionObject); r[1] = code(o1->m_collisionObject);
        for (int k = 0; k < 3; k++) { r[2 + k] = cp.m_normalWorldOnB[k]; r[5 + k] = cp.m_positionWorldOnB[k]; }
        r[8] = cp.m_distance1; r[9] = (float)idx0; r[10] = (float)idx1; r[11] = (float)part1;
    }
    return g_prev_added ? g_prev_added(cp, o0, part0, idx0, o1, part1, idx1) : true;
}
void ref_trace_contacts(int enable) {
    if (enable && gContactAddedCallback != trace_added) { g_prev_added = gContactAddedCallback; gContactAddedCallback = trace_added; }
    if (!enable && gContactAddedCallback == trace_added) gContactAddedCallback = g_prev_added;
    g_trace_n = 0;
}
int ref_trace_read(float* out, int maxRows) {
    int n = g_trace_n < maxRows ? g_trace_n : maxRows;
    memcpy(out, g_trace, sizeof(float) * 12 * n);
    g_trace_n = 0;
    return n;
}

// ---- narrowphase probes (unit parity of rl_gjk.h / rl_epa.h / rl_boxbox.h against the reference's own detectors) ----
// broadphase unique ids of the ball and the cars (pair body order: btHashedOverlappingPairCache::internalAddPair)
void ref_arena_unique_ids(void* h, int32_t* out) {
    Arena* a = (Arena*)h;
    out[0] = a->ball->_rigidBody.getBroadphaseHandle()->m_uniqueId;
    int n = (int)a->_cars.size();
    for (int i = 0; i < n; i++) out[1 + i] = car_by_id(a, i + 1)->_rigidBody.getBroadphaseHandle()->m_uniqueId;
}

struct ProbeResult : public btDiscreteCollisionDetectorInterface::Result {
    int n = 0; btVector3 normal, point; btScalar depth = 0;
    void setShapeIdentifiersA(int, int) override {}
    void setShapeIdentifiersB(int, int) override {}
    void addContactPoint(const btVector3& normalOnBInWorld, const btVector3& pointInWorld, btScalar d) override {
        n++; normal = normalOnBInWorld; point = pointInWorld; depth = d;
    }
};

// btConvexConvexAlgorithm::processCollision's detector call for (box with margin 0.04, triangle with margin 0):
// boxT = {origin xyz, basis rows 9}, tri = 3 x xyz (world, BT units), breaking = manifold contact-breaking threshold.
// out = {normal xyz, point xyz, depth}; returns the number of contact points reported (0 / 1) and the detector's last method.
int ref_probe_box_triangle(const float* halfExt, const float* boxT, const float* tri, float breaking, float* out, int* method) {
    btBoxShape box(btVector3(halfExt[0], halfExt[1], halfExt[2]));
    btTriangleShape tm(btVector3(tri[0], tri[1], tri[2]), btVector3(tri[3], tri[4], tri[5]), btVector3(tri[6], tri[7], tri[8]));
    tm.setMargin(0.f);
    btVoronoiSimplexSolver simplex;
    btGjkEpaPenetrationDepthSolver pd;
    btGjkPairDetector det(&box, &tm, &simplex, &pd);
    btGjkPairDetector::ClosestPointInput in;
    in.m_maximumDistanceSquared = box.getMargin() + tm.getMargin() + breaking;
    in.m_maximumDistanceSquared *= in.m_maximumDistanceSquared;
    in.m_transformA.setOrigin(btVector3(boxT[0], boxT[1], boxT[2]));
    in.m_transformA.setBasis(btMatrix3x3(boxT[3], boxT[4], boxT[5], boxT[6], boxT[7], boxT[8], boxT[9], boxT[10], boxT[11]));
    in.m_transformB.setIdentity();
    ProbeResult r;
    det.getClosestPoints(in, r, false);
    if (method) *method = det.m_lastUsedMethod;
    if (r.n) { for (int k = 0; k < 3; k++) { out[k] = r.normal[k]; out[3 + k] = r.point[k]; } out[6] = r.depth; }
    return r.n;
}
int ref_probe_box_sphere(const float* halfExt, const float* boxT, const float* center, float radius, float breaking, float* out, int* method) {
    btBoxShape box(btVector3(halfExt[0], halfExt[1], halfExt[2]));
    btSphereShape sp(radius);
    btVoronoiSimplexSolver simplex;
    btGjkEpaPenetrationDepthSolver pd;
    btGjkPairDetector det(&box, &sp, &simplex, &pd);
    btGjkPairDetector::ClosestPointInput in;
    in.m_maximumDistanceSquared = box.getMargin() + sp.getMargin() + breaking;
    in.m_maximumDistanceSquared *= in.m_maximumDistanceSquared;
    in.m_transformA.setOrigin(btVector3(boxT[0], boxT[1], boxT[2]));
    in.m_transformA.setBasis(btMatrix3x3(boxT[3], boxT[4], boxT[5], boxT[6], boxT[7], boxT[8], boxT[9], boxT[10], boxT[11]));
    in.m_transformB.setIdentity();
    in.m_transformB.setOrigin(btVector3(center[0], center[1], center[2]));
    ProbeResult r;
    det.getClosestPoints(in, r, false);
    if (method) *method = det.m_lastUsedMethod;
    if (r.n) { for (int k = 0; k < 3; k++) { out[k] = r.normal[k]; out[3 + k] = r.point[k]; } out[6] = r.depth; }
    return r.n;
}
// btBoxBoxDetector for two boxes: out rows of {normal xyz, point xyz, depth}; returns the number of points
struct ProbeMulti : public btDiscreteCollisionDetectorInterface::Result {
    int n = 0; float* out; int cap;
    void setShapeIdentifiersA(int, int) override {}
    void setShapeIdentifiersB(int, int) override {}
    void addContactPoint(const btVector3& nrm, const btVector3& p, btScalar d) override {
        if (n < cap) { float* r = out + 7 * n; for (int k = 0; k < 3; k++) { r[k] = nrm[k]; r[3 + k] = p[k]; } r[6] = d; }
        n++;
    }
};
int ref_probe_box_box(const float* halfA, const float* tA, const float* halfB, const float* tB, float* out, int cap) {
    btBoxShape a(btVector3(halfA[0], halfA[1], halfA[2])), b(btVector3(halfB[0], halfB[1], halfB[2]));
    btBoxBoxDetector det(&a, &b);
    btDiscreteCollisionDetectorInterface::ClosestPointInput in;
    in.m_maximumDistanceSquared = BT_LARGE_FLOAT;
    in.m_transformA.setOrigin(btVector3(tA[0], tA[1], tA[2]));
    in.m_transformA.setBasis(btMatrix3x3(tA[3], tA[4], tA[5], tA[6], tA[7], tA[8], tA[9], tA[10], tA[11]));
    in.m_transformB.setOrigin(btVector3(tB[0], tB[1], tB[2]));
    in.m_transformB.setBasis(btMatrix3x3(tB[3], tB[4], tB[5], tB[6], tB[7], tB[8], tB[9], tB[10], tB[11]));
    ProbeMulti r; r.out = out; r.cap = cap;
    det.getClosestPoints(in, r, false);
    return r.n;
}

// iteration order of Arena::_cars (the player order of every gym-level vector)
void ref_arena_player_order(void* h, int32_t* ids) {
    int k = 0;
    for (Car* c : ((Arena*)h)->_cars) ids[k++] = (int32_t)c->id;
}

// ---- gym -------------------------------------------------------------------
void* ref_gym_create(const rlg_engine_cfg* cfg) {
    try { return make_gym(cfg); } catch (std::exception&) { return nullptr; }
}
void ref_gym_destroy(void* h) { delete (RefGym*)h; }
void* ref_gym_arena(void* h) { return ((RefGym*)h)->gym->arena; }
int ref_gym_num_players(void* h) { return ((RefGym*)h)->match->playerAmount; }

int ref_gym_reset(void* h, float* obs_out) {
    RefGym* g = (RefGym*)h;
    return flatten_obs(g->gym->Reset(), obs_out);
}

// Gym::Reset without the state setter: adopt whatever the arena currently holds
// (used after ref_arena_set_state to start an episode from an injected state).
int ref_gym_reset_from_current(void* h, float* obs_out) {
    RefGym* g = (RefGym*)h;
    GameState s(g->gym->arena);
    g->match->EpisodeReset(s);
    g->gym->prevState = s;
    g->gym->eventTracker.ResetPersistentInfo();
    return flatten_obs(g->match->BuildObservations(s), obs_out);
}

int ref_gym_step(void* h, const int32_t* actions, float* obs_out, float* rew_out, uint8_t* done_out) {
    RefGym* g = (RefGym*)h;
    IList acts(actions, actions + g->match->playerAmount);
    auto r = g->gym->Step(acts);
    int w = flatten_obs(r.obs, obs_out);
    for (size_t i = 0; i < r.reward.size(); i++) rew_out[i] = r.reward[i];
    *done_out = r.done;
    return w;
}

// Gym::Step (Gym.cpp:68-102) minus arena->Step and eventTracker.Update: evaluates the plugins on the arena's CURRENT state.
int ref_gym_eval_current(void* h, const int32_t* actions, float* obs_out, float* rew_out, uint8_t* done_out) {
    RefGym* g = (RefGym*)h;
    IList acts(actions, actions + g->match->playerAmount);
    ActionSet parsed = g->match->ParseActions(acts, g->gym->prevState);
    g->match->prevActions = parsed;
    GameState state = g->gym->prevState;
    state.UpdateFromArena(g->gym->arena);
    FList2 obs = g->match->BuildObservations(state);
    bool done = g->match->IsDone(state);
    FList rewards = g->match->GetRewards(state, done);
    g->gym->prevState = state;
    int w = flatten_obs(obs, obs_out);
    for (size_t i = 0; i < rewards.size(); i++) rew_out[i] = rewards[i];
    *done_out = done;
    return w;
}

// gym-level extras of the last GameState (what the fused gym kernel must reproduce)
void ref_gym_last_state(void* h, int32_t* score_line2, int32_t* last_touch, int32_t* match_counters /*[P*8]*/,
                        uint8_t* touched_step /*[P]*/) {
    RefGym* g = (RefGym*)h;
    const GameState& s = g->gym->prevState;
    score_line2[0] = s.scoreLine[0]; score_line2[1] = s.scoreLine[1];
    *last_touch = s.lastTouchCarID;
    for (size_t i = 0; i < s.players.size(); i++) {
        const PlayerData& p = s.players[i];
        int32_t* c = match_counters + i * 8;
        c[0] = p.matchGoals; c[1] = p.matchSaves; c[2] = p.matchAssists; c[3] = p.matchShots;
        c[4] = p.matchShotPasses; c[5] = p.matchBumps; c[6] = p.matchDemos; c[7] = p.boostPickups;
        touched_step[i] = p.ballTouchedStep;
    }
}

int ref_action_table(float* out) {
    DiscreteAction d;
    int k = 0;
    for (auto& a : d.actions) for (int i = 0; i < 8; i++) out[k++] = a[i];
    return (int)d.actions.size();
}

// ---- CPU baseline: the reference's multithreaded collection loop, sim only --
// T threads x G gyms, each thread looping Gym::Step + auto-reset exactly like
// GameInst::Step (src/public/RLGymPPO_CPP/Threading/GameInst.cpp:7-38), actions =
// uniform random index in [0,90). Returns player-steps per second.
double ref_bench_collect(const rlg_engine_cfg* cfg, int num_threads, int gyms_per_thread,
                         int warmup_steps, int timed_steps, uint32_t seed) {
    std::vector<std::vector<RefGym*>> gyms(num_threads);
    for (int t = 0; t < num_threads; t++)
        for (int i = 0; i < gyms_per_thread; i++) gyms[t].push_back(make_gym(cfg));
    std::atomic<int> ready{0}; std::atomic<bool> go{false};
    std::vector<double> secs(num_threads, 0.0);
    std::vector<std::thread> ths;
    int P = gyms[0][0]->match->playerAmount;
    for (int t = 0; t < num_threads; t++) {
        ths.emplace_back([&, t] {
            RocketSim::Math::GetRandEngine().seed(seed + 977 * t);
            uint32_t rng = seed * 2654435761u + t * 40503u + 1;
            auto next = [&rng] { rng ^= rng << 13; rng ^= rng >> 17; rng ^= rng << 5; return rng; };
            for (auto g : gyms[t]) g->gym->Reset();
            IList acts(P);
            auto run = [&](int steps) {
                for (int s = 0; s < steps; s++)
                    for (auto g : gyms[t]) {
                        for (int p = 0; p < P; p++) acts[p] = next() % 90;
                        auto r = g->gym->Step(acts);
                        if (r.done) g->gym->Reset();
                    }
            };
            run(warmup_steps);
            ready++;
            while (!go.load()) std::this_thread::yield();
            auto t0 = std::chrono::steady_clock::now();
            run(timed_steps);
            secs[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        });
    }
    while (ready.load() < num_threads) std::this_thread::yield();
    go = true;
    for (auto& th : ths) th.join();
    double mx = 0;
    for (double s : secs) mx = std::max(mx, s);
    for (auto& v : gyms) for (auto g : v) delete g;
    double total = (double)num_threads * gyms_per_thread * timed_steps * P;
    return total / mx;
}

// persistent variant for bench.py --impl reference: gyms are built once, every call runs `steps` env-steps per gym
struct RefBench {
    std::vector<std::vector<RefGym*>> gyms;
    int P = 0;
    uint32_t seed = 1;
    uint64_t calls = 0;
};
void* ref_bench_create(const rlg_engine_cfg* cfg, int num_threads, int gyms_per_thread, uint32_t seed) {
    RefBench* b = new RefBench();
    b->gyms.resize(num_threads);
    b->seed = seed;
    std::vector<std::thread> ths;
    for (int t = 0; t < num_threads; t++)
        ths.emplace_back([&, t] {
            RocketSim::Math::GetRandEngine().seed(seed + 977 * t);
            for (int i = 0; i < gyms_per_thread; i++) { RefGym* g = make_gym(cfg); g->gym->Reset(); b->gyms[t].push_back(g); }
        });
    for (auto& th : ths) th.join();
    b->P = b->gyms[0][0]->match->playerAmount;
    return b;
}
// returns seconds (max over threads) for `steps` Gym::Step per gym, GameInst::Step-style auto reset
double ref_bench_run(void* h, int steps) {
    RefBench* b = (RefBench*)h;
    int T = (int)b->gyms.size();
    std::atomic<int> ready{0}; std::atomic<bool> go{false};
    std::vector<double> secs(T, 0.0);
    std::vector<std::thread> ths;
    uint64_t call = b->calls++;
    for (int t = 0; t < T; t++) {
        ths.emplace_back([&, t] {
            RocketSim::Math::GetRandEngine().seed(b->seed + 977 * t + 31 * (uint32_t)call);
            uint32_t rng = (b->seed + (uint32_t)call * 7919u) * 2654435761u + t * 40503u + 1;
            auto next = [&rng] { rng ^= rng << 13; rng ^= rng >> 17; rng ^= rng << 5; return rng; };
            IList acts(b->P);
            ready++;
            while (!go.load()) std::this_thread::yield();
            auto t0 = std::chrono::steady_clock::now();
            for (int s = 0; s < steps; s++)
                for (auto g : b->gyms[t]) {
                    for (int p = 0; p < b->P; p++) acts[p] = next() % 90;
                    auto r = g->gym->Step(acts);
                    if (r.done) g->gym->Reset();
                }
            secs[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        });
    }
    while (ready.load() < T) std::this_thread::yield();
    go = true;
    for (auto& th : ths) th.join();
    double mx = 0;
    for (double s : secs) mx = std::max(mx, s);
    return mx;
}
void ref_bench_destroy(void* h) {
    RefBench* b = (RefBench*)h;
    for (auto& v : b->gyms) for (auto g : v) delete g;
    delete b;
}

// ---- policy-driven collection for the learning-curve A/B (tools/learning_curves.py): persistent worker threads, one call steps every
// gym once with the caller's actions, GameInst::Step semantics (GameInst.cpp:7-38: a finished gym is reset and hands back the reset obs)
struct RefPool {
    std::vector<std::vector<RefGym*>> gyms;
    std::vector<std::thread> workers;
    int P = 0, obs = 0, G = 0;
    std::atomic<uint64_t> generation{0};
    std::atomic<int> finished{0};
    std::atomic<bool> quit{false};
    const int32_t* actions = nullptr; float* obsOut = nullptr; float* rewOut = nullptr; uint8_t* doneOut = nullptr;
    int mode = 0;  // 0 = step, 1 = reset all + obs
};
static void pool_worker(RefPool* b, int t, uint32_t seed) {
    RocketSim::Math::GetRandEngine().seed(seed + 977 * t);
    uint64_t seen = 0;
    const int perThread = (int)b->gyms[t].size();
    while (true) {
        while (b->generation.load(std::memory_order_acquire) == seen) { if (b->quit.load()) return; std::this_thread::yield(); }
        seen = b->generation.load(std::memory_order_acquire);
        int base = 0;
        for (int k = 0; k < t; k++) base += (int)b->gyms[k].size();
        for (int i = 0; i < perThread; i++) {
            RefGym* g = b->gyms[t][i];
            const int gi = base + i;
            float* o = b->obsOut + (size_t)gi * b->P * b->obs;
            if (b->mode == 1) { flatten_obs(g->gym->Reset(), o); continue; }
            IList acts(b->actions + (size_t)gi * b->P, b->actions + (size_t)(gi + 1) * b->P);
            auto r = g->gym->Step(acts);
            for (int p = 0; p < b->P; p++) b->rewOut[(size_t)gi * b->P + p] = r.reward[p];
            b->doneOut[gi] = r.done;
            if (r.done) flatten_obs(g->gym->Reset(), o); else flatten_obs(r.obs, o);
        }
        b->finished.fetch_add(1, std::memory_order_release);
    }
}
void* ref_pool_create(const rlg_engine_cfg* cfg, int num_threads, int gyms_per_thread, uint32_t seed) {
    RefPool* b = new RefPool();
    b->gyms.resize(num_threads);
    {
        std::vector<std::thread> ths;
        for (int t = 0; t < num_threads; t++)
            ths.emplace_back([&, t] {
                RocketSim::Math::GetRandEngine().seed(seed + 977 * t);
                for (int i = 0; i < gyms_per_thread; i++) b->gyms[t].push_back(make_gym(cfg));
            });
        for (auto& th : ths) th.join();
    }
    b->P = b->gyms[0][0]->match->playerAmount;
    b->G = num_threads * gyms_per_thread;
    std::vector<float> tmp(4096);
    b->obs = flatten_obs(b->gyms[0][0]->gym->Reset(), tmp.data()) ;
    for (int t = 0; t < num_threads; t++) b->workers.emplace_back(pool_worker, b, t, seed + 13);
    return b;
}
int ref_pool_obs_size(void* h) { return ((RefPool*)h)->obs; }
static void pool_run(RefPool* b, int mode) {
    b->mode = mode;
    b->finished.store(0);
    b->generation.fetch_add(1, std::memory_order_release);
    while (b->finished.load(std::memory_order_acquire) < (int)b->workers.size()) std::this_thread::yield();
}
void ref_pool_reset(void* h, float* obs_out) { RefPool* b = (RefPool*)h; b->obsOut = obs_out; pool_run(b, 1); }
void ref_pool_step(void* h, const int32_t* actions, float* obs_out, float* rew_out, uint8_t* done_out) {
    RefPool* b = (RefPool*)h;
    b->actions = actions; b->obsOut = obs_out; b->rewOut = rew_out; b->doneOut = done_out;
    pool_run(b, 0);
}
void ref_pool_destroy(void* h) {
    RefPool* b = (RefPool*)h;
    b->quit = true;
    for (auto& th : b->workers) th.join();
    for (auto& v : b->gyms) for (auto g : v) delete g;
    delete b;
}

size_t ref_sizeof_car_state() { return sizeof(rlg_car_state); }

} // extern "C"
