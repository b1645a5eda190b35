// rl_tick.h — Arena::Step for one tick (R/Sim/Arena/Arena.cpp:716-812) including the Bullet
// world step (B/BulletDynamics/Dynamics/btDiscreteDynamicsWorld.cpp:325-437), boost pads
// (R/Sim/BoostPad/BoostPad.cpp:51-108, BoostPadGrid.cpp:5-25) and, on top, Gym::Step (G/Gym.cpp:68-102).
#pragma once
#include "rl_solver.h"
#include "rl_gym.h"

namespace rl {

// ---- boost pads --------------------------------------------------------------------------------
RL_HDI bool pad_is_big(int i) { return i < kNumPadsBig; }

RL_HD inline void pads_pre_tick(ArenaS& a) {
    for (int i = 0; i < kNumPads; i++) {
        PadS& p = a.pads[i];
        if (p.cooldown > 0) p.cooldown = fmaxf_(p.cooldown - kTickTime, 0.f);
        p.isActive = (p.cooldown == 0);
    }
}

// BoostPadGrid::CheckCollision + BoostPad::_CheckCollide; locked[] = curLockedCar (car index + 1, 0 none)
RL_HD inline void pads_check_car(const ArenaS& a, const Tables& tb, const CarConsts& k, int ci, int32_t* locked) {
    const CarS& c = a.cars[ci];
    if (c.isDemoed || c.boost >= 100) return;
    V3 carPos = c.pos * BT2UU;
    const float EXTENT_Z = C::PAD_CYL_HEIGHT + 250.f;
    if (carPos.z > EXTENT_Z) return;
    const int CELLS_X = 8, CELLS_Y = 10;
    const int CELL_SIZE_X = (int)(4096.f / (CELLS_X / 2)), CELL_SIZE_Y = (int)(5120.f / (CELLS_Y / 2));
    int indexX = (int)(carPos.x / CELL_SIZE_X + (CELLS_X / 2));
    int indexY = (int)(carPos.y / CELL_SIZE_Y + (CELLS_Y / 2));
    for (int i = 0; i < kNumPads; i++) {
        V3 pp(tb.padPos[i * 3 + 0], tb.padPos[i * 3 + 1], tb.padPos[i * 3 + 2]);
        int px = (int)(pp.x / CELL_SIZE_X + (CELLS_X / 2));
        int py = (int)(pp.y / CELL_SIZE_Y + (CELLS_Y / 2));
        int lox = indexX - 1 > 0 ? indexX - 1 : 0, hix = indexX + 1 < CELLS_X - 1 ? indexX + 1 : CELLS_X - 1;
        int loy = indexY - 1 > 0 ? indexY - 1 : 0, hiy = indexY + 1 < CELLS_Y - 1 ? indexY + 1 : CELLS_Y - 1;
        if (px < lox || px > hix || py < loy || py > hiy) continue;
        V3 posBT = pp * UU2BT;
        bool big = pad_is_big(i);
        bool colliding = false;
        if (a.pads[i].prevLockedCarId == ci + 1) {
            float boxRad = (big ? C::PAD_BOX_RAD_BIG : C::PAD_BOX_RAD_SMALL) * UU2BT;
            V3 boxMin = posBT - V3(boxRad, boxRad, 0), boxMax = posBT + V3(boxRad, boxRad, C::PAD_BOX_HEIGHT * UU2BT);
            // car AABB: compound -> child box AABB
            V3 center = c.pos + c.rot * k.hitboxOffset;
            V3 ext(dot(vabs(c.rot.r[0]), k.halfExt), dot(vabs(c.rot.r[1]), k.halfExt), dot(vabs(c.rot.r[2]), k.halfExt));
            V3 cmn = center - ext, cmx = center + ext;
            colliding = (boxMax.x > cmn.x && boxMax.y > cmn.y && boxMax.z > cmn.z) && (boxMin.x < cmx.x && boxMin.y < cmx.y && boxMin.z < cmx.z);
        } else {
            float rad = (big ? C::PAD_CYL_RAD_BIG : C::PAD_CYL_RAD_SMALL) * UU2BT;
            float dx = c.pos.x - posBT.x, dy = c.pos.y - posBT.y;
            if (dx * dx + dy * dy < rad * rad) colliding = fabsf(c.pos.z - posBT.z) < (C::PAD_CYL_HEIGHT * UU2BT);
        }
        if (colliding) locked[i] = ci + 1;
    }
}

RL_HD inline void pads_post_tick(ArenaS& a, const int32_t* locked) {
    for (int i = 0; i < kNumPads; i++) {
        PadS& p = a.pads[i];
        int lockedId = 0;
        if (locked[i]) {
            lockedId = locked[i];
            if (p.isActive) {
                CarS& c = a.cars[locked[i] - 1];
                float add = pad_is_big(i) ? C::PAD_BOOST_BIG : C::PAD_BOOST_SMALL;
                c.boost = fminf_(c.boost + add, C::BOOST_MAX);
                p.isActive = 0;
                p.cooldown = pad_is_big(i) ? C::PAD_COOLDOWN_BIG : C::PAD_COOLDOWN_SMALL;
            }
        }
        p.prevLockedCarId = lockedId;
    }
}

#ifdef RL_DEBUG_CONTACTS
static ContactSet g_dbg_contacts;
#endif

// ---- one physics tick -----------------------------------------------------------------------------
RL_HD RL_NOINLINE void arena_tick(ArenaS& a, const SimCfg& cfg, const MeshSet& ms, const Tables& tb, int firstTickOfStep) {
    const float dt = kTickTime;
    const CarConsts k = car_consts();
    const Thresholds thr = contact_thresholds(k);
    TickW tw;
    tw.ballVelCache = V3(); tw.ballForce = V3();
    int64_t tick = get_i64(a.tickLo, a.tickHi);
    const int P = cfg.numCars;

    // ball zero-velocity sleeping (Arena.cpp:721-727)
    bool ballActive = !(len2(a.ball.vel) == 0.f && len2(a.ball.angvel) == 0.f);

    // activation state / contact response are decided at the top of Car::_PreTickUpdate, before a possible respawn,
    // and a car demolished DURING this tick still responds and integrates until the next tick (Car.cpp:38-41,69-87)
    int32_t noResponse[kMaxCars];
    for (int c = 0; c < P; c++) noResponse[c] = a.cars[c].isDemoed;
    for (int p = 0; p < P; p++) { int ci = cfg.playerOrder[p]; car_pre_tick(a, cfg, ms, k, ci, tw.cars[ci]); }
    if (P > 0) pads_pre_tick(a);

    // ---- btDiscreteDynamicsWorld::stepSimulation ----
    // applyGravity on active bodies
    const V3 g(0.f * UU2BT, 0.f * UU2BT, C::GRAVITY_Z * UU2BT);
    if (ballActive) tw.ballForce += g * C::BALL_MASS;
    for (int c = 0; c < P; c++) if (!noResponse[c]) tw.cars[c].force += g * C::CAR_MASS;
    // predictUnconstraintMotion: damping (ball only: linear 0.03)
    a.ball.vel = a.ball.vel * cfg.ballDampFactor;

    // collision detection, in the reference's pair order (btRSBroadphase::calculateOverlappingPairs)
    ContactSet cs; cs.n = 0; cs.overflow = 0;
    CollideCtx cx; cx.a = &a; cx.cfg = &cfg; cx.tw = &tw; cx.k = &k; cx.tick = tick; cx.firstTickOfStep = firstTickOfStep;
    for (int c = 0; c < kMaxCars; c++) cx.noResponse[c] = c < P ? noResponse[c] : 1;
    float ballR = C::BALL_RADIUS * UU2BT;
    float ballAabb = ballR + 0.08f;
    // a sleeping ball vs the (always "sleeping") static bodies is skipped by btCollisionDispatcher::needsCollision
    // (both inactive): on the tick it is woken by a car it has no world contacts yet.
    if (ballActive) {
        sphere_meshes(cx, cs, ms, a.ball.pos, ballR, thr.ball);
        for (int p = 0; p < 4; p++) sphere_plane(cx, cs, a.ball.pos, ballR, p, thr.ball);
    }
    bool ballWoken = false;
    V3 bmn = a.ball.pos - V3(ballAabb, ballAabb, ballAabb), bmx = a.ball.pos + V3(ballAabb, ballAabb, ballAabb);
    V3 cmn[kMaxCars], cmx[kMaxCars];
    for (int c = 0; c < P; c++) {
        const CarS& car = a.cars[c];
        V3 center = car.pos + car.rot * k.hitboxOffset;
        V3 ext(dot(vabs(car.rot.r[0]), k.halfExt), dot(vabs(car.rot.r[1]), k.halfExt), dot(vabs(car.rot.r[2]), k.halfExt));
        cmn[c] = center - ext; cmx[c] = center + ext;
    }
    auto overlap = [](V3 amn, V3 amx, V3 bmn_, V3 bmx_) {
        return !(amn.x > bmx_.x || amx.x < bmn_.x || amn.y > bmx_.y || amx.y < bmn_.y || amn.z > bmx_.z || amx.z < bmn_.z);
    };
    float thrCarBall = fminf_(thr.ball, thr.car);
    for (int c = 0; c < P; c++) {
        if (!overlap(bmn, bmx, cmn[c], cmx[c])) continue;
        if (!noResponse[c]) ballWoken = true;  // islands merge on broadphase overlap
        if (!ballActive && noResponse[c]) continue;
        car_ball(cx, cs, c, thrCarBall);
    }
    for (int c = 0; c < P; c++) {
        box_meshes(cx, cs, ms, c, thr.car);
        for (int p = 0; p < 4; p++) box_plane(cx, cs, c, p, thr.car);
        for (int d = c + 1; d < P; d++) {
            if (!overlap(cmn[c], cmx[c], cmn[d], cmx[d])) continue;
            car_car(cx, cs, c, d, thr.car);
        }
    }

#ifdef RL_DEBUG_CONTACTS
    g_dbg_contacts = cs;
#endif
    // ---- solve ----
    SolverBody sb[1 + kMaxCars];
    {
        SolverBody& b = sb[0];
        b.pos = a.ball.pos; b.rot = M3::identity();
        b.linVel = a.ball.vel; b.angVel = a.ball.angvel;
        b.invMass = 1.f / C::BALL_MASS;
        float inertia = 0.4f * C::BALL_MASS * ballR * ballR;
        float ii = 1.f / inertia;
        b.invInertiaWorld = M3(V3(ii, 0, 0), V3(0, ii, 0), V3(0, 0, ii));
        b.extForceImp = tw.ballForce * b.invMass * dt;
        b.extTorqueImp = V3();
        b.dLin = b.dAng = b.push = b.turn = V3();
        b.active = ballActive || ballWoken;
    }
    for (int c = 0; c < P; c++) {
        SolverBody& b = sb[1 + c];
        const CarS& car = a.cars[c];
        b.pos = car.pos; b.rot = car.rot; b.linVel = car.vel; b.angVel = car.angvel;
        b.invMass = k.invMass;
        b.invInertiaWorld = tw.cars[c].invInertiaWorld;
        b.extForceImp = tw.cars[c].force * b.invMass * dt;
        b.extTorqueImp = tmul(tw.cars[c].torque, b.invInertiaWorld) * dt;
        b.dLin = b.dAng = b.push = b.turn = V3();
        b.active = !noResponse[c];
    }
    solve_arena(sb, 1 + P, cs);

    // ---- integrateTransforms ----
    if (sb[0].active) {
        a.ball.vel = sb[0].linVel; a.ball.angvel = sb[0].angVel;
        a.ball.pos = sb[0].pos + a.ball.vel * dt;  // integrateTransformNoRot
    }
    for (int c = 0; c < P; c++) {
        if (!sb[1 + c].active) continue;
        CarS& car = a.cars[c];
        car.vel = sb[1 + c].linVel; car.angvel = sb[1 + c].angVel;
        car.pos = sb[1 + c].pos; car.rot = sb[1 + c].rot;
        integrate_transform(car.pos, car.rot, car.vel, car.angvel, dt);
    }

    // ---- post tick ----
    int32_t locked[kNumPads];
    for (int i = 0; i < kNumPads; i++) locked[i] = 0;
    for (int p = 0; p < P; p++) {
        int ci = cfg.playerOrder[p];
        car_post_tick(a.cars[ci], tw.cars[ci]);
        pads_check_car(a, tb, k, ci, locked);
    }
    if (P > 0) pads_post_tick(a, locked);
    // Ball::_FinishPhysicsTick (Ball.cpp:112-138)
    if (!is_zero(tw.ballVelCache)) a.ball.vel += tw.ballVelCache;
    {
        const float maxSpeed = C::BALL_MAX_SPEED * UU2BT;
        if (len2(a.ball.vel) > maxSpeed * maxSpeed) a.ball.vel = normalized(a.ball.vel) * maxSpeed;
        if (len2(a.ball.angvel) > C::BALL_MAX_ANG_SPEED * C::BALL_MAX_ANG_SPEED) a.ball.angvel = normalized(a.ball.angvel) * C::BALL_MAX_ANG_SPEED;
    }
    a.ball.updateCounterLo++;
    set_i64(a.tickLo, a.tickHi, tick + 1);
}

// ---- Gym::Step (G/Gym.cpp:68-102) + GameInst::Step auto-reset --------------------------------------
RL_HD inline void gym_step(ArenaS& a, const SimCfg& cfg, const MeshSet& ms, const Tables& tb, const int32_t* actionIdx,
                           float* obsOut, float* rewardOut, uint8_t* doneOut) {
    parse_actions(a, cfg, tb, actionIdx);
    arena_tick(a, cfg, ms, tb, 1);
    event_tracker_update(a, cfg);
    snapshot_update(a, cfg);
    build_obs(a, cfg, tb, obsOut);
    bool done = compute_done(a, cfg);
    compute_rewards(a, cfg, rewardOut);
    *doneOut = done ? 1 : 0;
    for (int t = 1; t < cfg.tickSkip; t++) arena_tick(a, cfg, ms, tb, 0);
    if (done) {
        gym_reset(a, cfg);
        build_obs(a, cfg, tb, obsOut);
    }
}

}  // namespace rl
