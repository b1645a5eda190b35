// rl_tick.h — Arena::Step for one tick (R/Sim/Arena/Arena.cpp:716-812) including the Bullet
// world step (B/BulletDynamics/Dynamics/btDiscreteDynamicsWorld.cpp:325-437), boost pads
// (R/Sim/BoostPad/BoostPad.cpp:51-108, BoostPadGrid.cpp:5-25) and, on top, Gym::Step (G/Gym.cpp:68-102).
#pragma once
#include "rl_solver.h"
#include "rl_gym.h"

namespace rl {

// ---- boost pads --------------------------------------------------------------------------------
RL_HDI bool pad_is_big(int i) { return i < kNumPadsBig; }

// BoostPad::_PreTickUpdate for all pads (BoostPad.cpp:51-58): only pads with cooldown > 0 are touched, then
// isActive = (cooldown == 0) for every pad
RL_HD inline void pads_pre_tick(ArenaS& a) {
    uint64_t cooling = pads_cooling(a.pads);
    for (uint64_t m = cooling; m; m &= m - 1) {
        int i = lowest_bit(m);
        float cd = fmaxf_(a.pads.cooldown[i] - kTickTime, 0.f);
        a.pads.cooldown[i] = cd;
        if (cd == 0.f) cooling &= ~(1ULL << i);
    }
    pads_set_cooling(a.pads, cooling);
    pads_set_active(a.pads, ~cooling & kAllPadsMask);
}

// BoostPadGrid::CheckCollision + BoostPad::_CheckCollide (BoostPadGrid.cpp:5-25, BoostPad.cpp:60-92) for one car:
// returns the mask of pads the car is colliding with this tick. The 3x3 cell neighbourhood is a table lookup
// (Tables::padCellMask, built with the reference's index rule).
RL_HD inline uint64_t pads_check_car(const ArenaS& a, const Tables& tb, const CarConsts& k, int ci) {
    const CarS& c = a.cars[ci];
    if (c.isDemoed || c.boost >= 100) return 0;
    V3 carPos = c.pos * BT2UU;
    const float EXTENT_Z = C::PAD_CYL_HEIGHT + 250.f;
    if (carPos.z > EXTENT_Z) return 0;
    const int CELLS_X = 8, CELLS_Y = 10;
    const int CELL_SIZE_X = (int)(4096.f / (CELLS_X / 2)), CELL_SIZE_Y = (int)(5120.f / (CELLS_Y / 2));
    int indexX = (int)(carPos.x / CELL_SIZE_X + (CELLS_X / 2));
    int indexY = (int)(carPos.y / CELL_SIZE_Y + (CELLS_Y / 2));
    if (indexX < -1 || indexX > CELLS_X || indexY < -1 || indexY > CELLS_Y) return 0;
    int cell = (indexX + 1) + (CELLS_X + 2) * (indexY + 1);
    uint64_t cand = ((uint64_t)tb.padCellMask[cell * 2 + 1] << 32) | tb.padCellMask[cell * 2];
    uint64_t hit = 0;
    for (uint64_t m = cand; m; m &= m - 1) {
        int i = lowest_bit(m);
        V3 posBT(tb.padPosBT[i * 3 + 0], tb.padPosBT[i * 3 + 1], tb.padPosBT[i * 3 + 2]);
        bool big = pad_is_big(i);
        bool colliding = false;
        if (pad_locked(a.pads, i) == ci + 1) {
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
        if (colliding) hit |= 1ULL << i;
    }
    return hit;
}

// BoostPad::_PostTickUpdate (BoostPad.cpp:94-108): hitMask[ci] from pads_check_car; the LAST car in _cars order that
// collides with a pad locks it (BoostPad.cpp:88 overwrites _internalState.curLockedCar)
RL_HD inline void pads_post_tick(ArenaS& a, const SimCfg& cfg, const uint64_t* hitMask) {
    uint64_t any = 0;
    for (int c = 0; c < cfg.numCars; c++) any |= hitMask[c];
    for (int i = 0; i < (kNumPads + 3) / 4; i++) a.pads.locked[i] = 0u;
    if (!any) return;
    uint64_t active = pads_active(a.pads), cooling = pads_cooling(a.pads);
    for (uint64_t m = any; m; m &= m - 1) {
        int i = lowest_bit(m);
        int lockedCar = -1;
        for (int p = 0; p < cfg.numCars; p++) { int ci = cfg.playerOrder[p]; if ((hitMask[ci] >> i) & 1ULL) lockedCar = ci; }
        if ((active >> i) & 1ULL) {
            CarS& c = a.cars[lockedCar];
            float add = pad_is_big(i) ? C::PAD_BOOST_BIG : C::PAD_BOOST_SMALL;
            c.boost = fminf_(c.boost + add, C::BOOST_MAX);
            active &= ~(1ULL << i);
            a.pads.cooldown[i] = pad_is_big(i) ? cfg.mut.padCooldownBig : cfg.mut.padCooldownSmall;
            cooling |= 1ULL << i;
        }
        pad_set_locked(a.pads, i, lockedCar + 1);
    }
    pads_set_active(a.pads, active); pads_set_cooling(a.pads, cooling);
}

#ifdef RL_DEBUG_CONTACTS
static ContactSet g_dbg_contacts;
static CarW g_dbg_carw[kMaxCars];  // the wheels' workspace after the last tick (host debugging only)
static int g_dbg_nmesh[kMaxCars], g_dbg_nplane[kMaxCars];  // a car's world-contact piece sizes as P1 left them (the mesh count's word is reused by P3)
#endif

// ---- one physics tick, split by ROLE ------------------------------------------------------------------------------
// Arena::Step (R/Sim/Arena/Arena.cpp:716-812) + btDiscreteDynamicsWorld::internalSingleStepSimulation
// (B/BulletDynamics/Dynamics/btDiscreteDynamicsWorld.cpp:393-437) for one arena, executed by 1 + numCars roles:
// role 0 = ball / arena, role 1+c = car c.  On the device each role of an arena is one lane of a different warp of the
// same block and the phases are separated by __syncthreads() (engine.cu k_roles); the host test build runs the same
// phases role after role (arena_tick).  Phase order and who-writes-what:
//
//   S0  every role snapshots its own body into the exchange (TickX)                                  | barrier B1
//   P1  car c : Car::_PreTickUpdate (wheel rays read the snapshots), gravity, hitbox AABB,
//               car-ball, hitbox-mesh, hitbox-plane narrowphase -> own contact segments
//       ball  : pads pre-tick, sphere-mesh, sphere-plane narrowphase -> own contact segment          | barrier B2
//       (role kernel: P1 is split in two by one more barrier — P1a: the cars find their mesh candidates, run the wheel rays'
//        mesh part and publish the hitbox pre-filter while the ball does its own narrowphase; P1b: the cars run the vehicle /
//        control model, car-ball and hitbox-plane while the otherwise idle BALL warp evaluates the cars' hitbox-mesh pairs,
//        which only depend on the start-of-tick pose; engine.cu)
//   P2  ball  : ball damping, car-car pairs (+bump/demo callbacks), sequential-impulse solve of the island(s) of the
//               ball and the COUPLED cars (contacts gathered in the reference's manifold order), write back,
//               integrate + finish the ball, tick count
//       car c : if uncoupled (no car-ball / car-car manifold possible): solve its own island, then P3   | barrier B3
//   P3  car c : (coupled cars; the others did it before B3) integrate transform,
//               Car::_PostTickUpdate/_FinishPhysicsTick, boost-pad overlap mask                          | barrier B4
//   P4  ball  : BoostPad::_PostTickUpdate (pick-ups)
struct Thresholds;

// segment pointers inside one arena's contact scratch
RL_HDI Contact* seg_ball(Contact* scratch) { return scratch; }
RL_HDI Contact* seg_car(Contact* scratch, int c) { return scratch + kSegBall + c * kSegCar; }
RL_HDI Contact* seg_pair(Contact* scratch, int ncars) { return scratch + kSegBall + ncars * kSegCar; }
RL_HDI Contact* seg_car_world(Contact* scratch, int c) { return seg_car(scratch, c) + 1; }                  // hitbox-mesh contacts, then ...
RL_HDI Contact* seg_car_plane(Contact* scratch, int c) { return seg_car(scratch, c) + 1 + kSegCarWorld; }   // ... the staged hitbox-plane contacts
// a car's world contacts as the solver sees them: the mesh piece followed by the plane piece, kSegCarWorld at most
RL_HDI int car_plane_taken(const CarX& o) { int room = kSegCarWorld - car_mesh_count(o); return o.nCarPlane < room ? o.nCarPlane : (room > 0 ? room : 0); }
RL_HDI int car_world_count(const CarX& o) { return car_mesh_count(o) + car_plane_taken(o); }

RL_HDI uint32_t respawn_rnd(const ArenaS& a, int ci) {
    uint64_t z = (((uint64_t)a.rngHi << 32) | a.rngLo) + 0x9E3779B97F4A7C15ULL * (uint64_t)(uint32_t)(a.tickLo + 1) + 0xD1B54A32D192ED03ULL * (uint64_t)(ci + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}

RL_HDI void tick_s0_car(const ArenaS& a, TickX x, int c) {
    const CarS& car = a.cars[c];
    CarX& o = x.car[c];
    o.pos = car.pos; o.vel = car.vel; o.angvel = car.angvel; o.rot = car.rot; o.demoed = car.isDemoed;
}
RL_HDI void tick_s0_ball(const ArenaS& a, TickX x) {
    x.h->ballPos = a.ball.pos; x.h->ballVel = a.ball.vel; x.h->ballAngvel = a.ball.angvel;
    // ball zero-velocity sleeping (Arena.cpp:721-727)
    x.h->ballActive = !(len2(a.ball.vel) == 0.f && len2(a.ball.angvel) == 0.f);
}

// tick_p1_car = pose (respawn, mesh candidates) + mesh part of the wheel rays + begin (vehicle update, control model,
// car-ball) + hitbox-mesh narrowphase + end (hitbox-plane, counts).  The role kernel replaces the two mesh parts by
// warp-cooperative passes (engine.cu cands_pass_warp / box_meshes_warp): same triangles, same order, same arithmetic
// per triangle (ray_tri4, box_mesh_item), so hits and contacts are identical.
RL_HD inline void tick_p1_car_pose(ArenaS& a, TickX x, const SimCfg& cfg, const MeshSet& ms, const CarConsts& k, int c, CarW& w, bool collect = true) {
    CarS& car = a.cars[c];
    CarX& o = x.car[c];
    // activation state / contact response are decided at the top of Car::_PreTickUpdate, before a possible respawn,
    // and a car demolished DURING this tick still responds and integrates until the next tick (Car.cpp:38-41,69-87)
    o.noResponse = car.isDemoed;
    o.ballVelCache = V3(); o.velCache = V3();
    w.solverDt = (a.tickLo | a.tickHi) ? kTickTime : (1.f / 60.f);
    RL_PT(-1);
    car_pre_tick_a(car, cfg, ms, k, c, w, respawn_rnd(a, c), collect);
}
// between the two: w.meshHit (+ the hitbox pre-filter) from wheel_mesh_rays or the role kernel's cooperative pass
RL_HD inline void tick_p1_car_begin(ArenaS& a, TickX x, const SimCfg& cfg, const MeshSet& ms, const CarConsts& k, const Thresholds& thr, int c, CarW& w,
                                    Contact* scratch, int firstTickOfStep, CollideCtx& cx, ContactSink& cw, const EpaCtx* epa = nullptr) {
    CarS& car = a.cars[c];
    CarX& o = x.car[c];
    car_pre_tick_b(car, x, cfg, ms, k, c, w);
    if (!o.noResponse) w.force += cfg.mut.gravityBT * C::CAR_MASS;  // applyGravity on active bodies
    o.force = w.force; o.torque = w.torque;
    V3 center = car.pos + car.rot * k.hitboxOffset;
    V3 ext(dot(vabs(car.rot.r[0]), k.halfExt), dot(vabs(car.rot.r[1]), k.halfExt), dot(vabs(car.rot.r[2]), k.halfExt));
    o.cmn = center - ext; o.cmx = center + ext;

    cx.a = &a; cx.cfg = &cfg; cx.tx = x; cx.k = &k; cx.tick = get_i64(a.tickLo, a.tickHi); cx.firstTickOfStep = firstTickOfStep; cx.epa = epa;
    cx.ballPos = x.h->ballPos; cx.ballVel = x.h->ballVel * cfg.ballDampFactor;  // predictUnconstraintMotion damping precedes the narrowphase
    // car-ball (manifold order: all car-ball pairs precede the car-world pairs)
    ContactSink cb = make_sink(seg_car(scratch, c), 1);
    float ballAabb = cfg.mut.ballRadius * UU2BT + 0.08f;
    V3 bmn = cx.ballPos - V3(ballAabb, ballAabb, ballAabb), bmx = cx.ballPos + V3(ballAabb, ballAabb, ballAabb);
    bool overlapBall = !(bmn.x > o.cmx.x || bmx.x < o.cmn.x || bmn.y > o.cmx.y || bmx.y < o.cmn.y || bmn.z > o.cmx.z || bmx.z < o.cmn.z);
    if (overlapBall && !(!x.h->ballActive && o.noResponse)) car_ball(cx, cb, c, fminf_(thr.ball, thr.car));
    o.nCarBall = cb.n;
    RL_PT(4);
    cw = make_sink(seg_car_world(scratch, c), kSegCarWorld);
}

// the hitbox-plane narrowphase into the staging slots (after the hitbox-mesh narrowphase in callback order: when both fire in one
// role the car's worldContact ends up the plane's; when the mesh piece ran elsewhere car_world_merge restores that order)
RL_HD inline void tick_p1_car_end(CollideCtx& cx, TickX x, const Thresholds& thr, int c, Contact* scratch) {
    RL_PT(5);
    ContactSink cpl = make_sink(seg_car_plane(scratch, c), kSegCarPlane);
#pragma unroll 1
    for (int p = 0; p < 4; p++) box_plane(cx, cpl, c, p, thr.car);
    x.car[c].nCarPlane = cpl.n;
    RL_PT(6);
}
// the car-world callback of a mesh piece that was evaluated by another role (CollideCtx::wcHas): it precedes the plane callbacks
// in the reference's manifold order, so it only sticks when no plane callback fired this tick
RL_HDI void car_world_merge(ArenaS& a, TickX x, int c, int meshHas, V3 meshNormal) {
    const CarX& o = x.car[c];
    if (meshHas && !(o.nCarPlane > 0 && !o.noResponse)) { a.cars[c].worldContactHas = 1; a.cars[c].worldContactNormal = meshNormal; }
}

RL_HD inline void tick_p1_car(ArenaS& a, TickX x, const SimCfg& cfg, const MeshSet& ms, const CarConsts& k, const Thresholds& thr, int c, CarW& w,
                              Contact* scratch, int firstTickOfStep) {
    CollideCtx cx; ContactSink cw;
    tick_p1_car_pose(a, x, cfg, ms, k, c, w);
    wheel_mesh_rays(a.cars[c], k, ms, w);
    tick_p1_car_begin(a, x, cfg, ms, k, thr, c, w, scratch, firstTickOfStep, cx, cw);
    if (w.cands.n >= 0) box_meshes_candidates(cx, cw, ms, w.cands, c, thr.car);
    else box_meshes(cx, cw, ms, c, thr.car);
    car_set_mesh_count(x.car[c], cw.n);
    tick_p1_car_end(cx, x, thr, c, scratch);
#if defined(RL_DEBUG_CONTACTS) && !defined(__CUDA_ARCH__)
    g_dbg_nmesh[c] = car_mesh_count(x.car[c]); g_dbg_nplane[c] = car_plane_taken(x.car[c]);
#endif
}

RL_HD inline void tick_p1_ball(ArenaS& a, TickX x, const SimCfg& cfg, const MeshSet& ms, const CarConsts& k, const Thresholds& thr, Contact* scratch) {
    RL_PT(-1);
    if (cfg.numCars > 0) pads_pre_tick(a);
    RL_PT(7);
    CollideCtx cx; cx.a = &a; cx.cfg = &cfg; cx.tx = x; cx.k = &k; cx.tick = get_i64(a.tickLo, a.tickHi); cx.firstTickOfStep = 0; cx.epa = nullptr;
    cx.ballPos = x.h->ballPos; cx.ballVel = x.h->ballVel * cfg.ballDampFactor;
    ContactSink cs = make_sink(seg_ball(scratch), kSegBall);
    // a sleeping ball vs the (always "sleeping") static bodies is skipped by btCollisionDispatcher::needsCollision
    // (both inactive): on the tick it is woken by a car it has no world contacts yet.
    if (x.h->ballActive) {
        float ballR = cfg.mut.ballRadius * UU2BT;
        sphere_meshes(cx, cs, ms, cx.ballPos, ballR, thr.ball);
        RL_PT(8);
#pragma unroll 1
        for (int p = 0; p < 4; p++) sphere_plane(cx, cs, cx.ballPos, ballR, p, thr.ball);
    }
    x.h->nBall = cs.n;
    RL_PT(9);
}

// ---- P2: constraint solve, one simulation island per role ---------------------------------------------------------
// A car is COUPLED this tick when it may share a manifold with another dynamic body: it produced a car-ball contact, or
// its hitbox AABB overlaps another car's (the broadphase condition for a car-car manifold).  Everything else it touches
// is static, so an uncoupled car is a simulation island of its own and its role solves it (tick_p2_car_self) while the
// ball role solves the island(s) made of the ball and the coupled cars (tick_p2_solve); see solve_island (rl_solver.h)
// for why this equals the reference's single solveGroup.
RL_HDI bool car_is_coupled(const TickX& x, int P, int c) {
    const CarX& o = x.car[c];
    if (o.noResponse) return false;  // not simulated this tick: nobody solves it
    if (o.nCarBall > 0) return true;
    for (int d = 0; d < P; d++) {
        if (d == c) continue;
        const CarX& od = x.car[d];
        bool ov = !(o.cmn.x > od.cmx.x || o.cmx.x < od.cmn.x || o.cmn.y > od.cmx.y || o.cmx.y < od.cmn.y || o.cmn.z > od.cmx.z || o.cmx.z < od.cmn.z);
        if (ov) return true;
    }
    return false;
}

RL_HDI void solver_body_from_car(SolverBody& b, const CarS& car, const CarX& o, const CarConsts& k) {
    const float dt = kTickTime;
    b.pos = car.pos; b.rot = car.rot; b.linVel = car.vel; b.angVel = car.angvel;
    b.invMass = k.invMass;
    b.invInertiaWorld = world_inertia(car.rot, k.invInertiaLocal);
    b.extForceImp = o.force * b.invMass * dt;
    b.extTorqueImp = tmul(o.torque, b.invInertiaWorld) * dt;
    b.dLin = b.dAng = b.push = b.turn = V3();
    b.active = !o.noResponse;
}

// an uncoupled car's own island: its car-world contacts only.  Returns false when the car is coupled (the ball role
// solves it and the car finishes its tick after the next barrier).
RL_HD inline bool tick_p2_car_self(ArenaS& a, TickX x, const SimCfg& cfg, const CarConsts& k, int c, Contact* scratch) {
    const CarX& o = x.car[c];
    RL_PT(-1);
    if (car_is_coupled(x, cfg.numCars, c)) return false;
    if (o.noResponse) return true;
    CarS& car = a.cars[c];
    const int nWorld = car_world_count(o);
    if (nWorld == 0) {
        // no manifold: the solver degenerates to writeBackBodies (btSequentialImpulseConstraintSolver.cpp:1878-1904):
        // v += 0 (deltas), then v += externalForceImpulse
        M3 iiw = world_inertia(car.rot, k.invInertiaLocal);
        car.vel = (car.vel + V3()) + o.force * k.invMass * kTickTime;
        car.angvel = (car.angvel + V3()) + tmul(o.torque, iiw) * kTickTime;
        return true;
    }
    SolverBody b;
    solver_body_from_car(b, car, o, k);
    {   // this car alone reads its world slots now: put the staged plane contacts behind the mesh contacts
        Contact* world = seg_car_world(scratch, c);
        const Contact* plane = seg_car_plane(scratch, c);
        const int nm = car_mesh_count(o), np = car_plane_taken(o);
        for (int i = 0; i < np; i++) world[nm + i] = plane[i];
    }
    solve_island_one(b, 1 + c, seg_car_world(scratch, c), nWorld);
    car.vel = b.linVel; car.angvel = b.angVel; car.pos = b.pos; car.rot = b.rot;
    RL_PT(13);
    return true;
}

RL_HD inline void tick_p2_solve(ArenaS& a, TickX x, const SimCfg& cfg, const CarConsts& k, const Thresholds& thr, Contact* scratch, int firstTickOfStep) {
    const float dt = kTickTime;
    const int P = cfg.numCars;
    const int64_t tick = get_i64(a.tickLo, a.tickHi);
    const bool ballActive = x.h->ballActive != 0;
    const float ballR = cfg.mut.ballRadius * UU2BT;
    RL_PT(-1);
    // predictUnconstraintMotion: damping (ball only: linear 0.03)
    a.ball.vel = a.ball.vel * cfg.ballDampFactor;

    // islands merge on broadphase overlap: a sleeping ball is woken by any responding car whose AABB overlaps
    bool ballWoken = false;
    uint32_t coupled = 0;
    {
        float ballAabb = ballR + 0.08f;
        V3 bmn = a.ball.pos - V3(ballAabb, ballAabb, ballAabb), bmx = a.ball.pos + V3(ballAabb, ballAabb, ballAabb);
        for (int c = 0; c < P; c++) {
            const CarX& o = x.car[c];
            bool ov = !(bmn.x > o.cmx.x || bmx.x < o.cmn.x || bmn.y > o.cmx.y || bmx.y < o.cmn.y || bmn.z > o.cmx.z || bmx.z < o.cmn.z);
            if (ov && !o.noResponse) ballWoken = true;
            if (car_is_coupled(x, P, c)) coupled |= 1u << c;
        }
    }
    // car-car pairs in the reference's pair order; contacts of pair (c, d > c) carry b == 1 + c, a == 1 + d (car_car)
    CollideCtx cx; cx.a = &a; cx.cfg = &cfg; cx.tx = x; cx.k = &k; cx.tick = tick; cx.firstTickOfStep = firstTickOfStep; cx.epa = nullptr;
    cx.ballPos = x.h->ballPos; cx.ballVel = a.ball.vel;
    ContactSink cp = make_sink(seg_pair(scratch, P), kSegPair);
    for (int c = 0; c < P; c++)
        for (int d = c + 1; d < P; d++) {
            const CarX &oc = x.car[c], &od = x.car[d];
            bool ov = !(oc.cmn.x > od.cmx.x || oc.cmx.x < od.cmn.x || oc.cmn.y > od.cmx.y || oc.cmx.y < od.cmn.y || oc.cmn.z > od.cmx.z || oc.cmx.z < od.cmn.z);
            if (ov) car_car(cx, cp, c, d, thr.car);
        }
    x.h->nPair = cp.n;
    RL_PT(10);

    const V3 gImp = cfg.mut.gravityBT * cfg.mut.ballMass;
#ifdef RL_DEBUG_CONTACTS
    {   // every contact of the tick in the reference's manifold order (host debugging only)
        ContactSet& cs = g_dbg_contacts; cs.n = 0; cs.overflow = 0;
        auto take = [&](const Contact* src, int n) { for (int i = 0; i < n; i++) { if (cs.n < kMaxContacts) cs.c[cs.n++] = src[i]; else cs.overflow++; } };
        take(seg_ball(scratch), x.h->nBall);
        for (int c = 0; c < P; c++) take(seg_car(scratch, c), x.car[c].nCarBall);
        for (int c = 0; c < P; c++) {
            // (an uncoupled car has already run P3 in the serial driver: its counts come from the P1 snapshot; its own island solve
            // moved the plane piece behind the mesh piece, so the world slots hold both)
            take(seg_car_world(scratch, c), g_dbg_nmesh[c]);
            take(((coupled >> c) & 1u) ? seg_car_plane(scratch, c) : seg_car_world(scratch, c) + g_dbg_nmesh[c], g_dbg_nplane[c]);
            const Contact* pr = seg_pair(scratch, P);
            for (int i = 0; i < cp.n; i++) if (pr[i].b == 1 + c) take(pr + i, 1);
        }
    }
#endif
    SolverBody sb[1 + kMaxCars];
    {
        SolverBody& b = sb[0];
        b.pos = a.ball.pos; b.rot = M3::identity();
        b.linVel = a.ball.vel; b.angVel = a.ball.angvel;
        b.invMass = 1.f / cfg.mut.ballMass;
        float inertia = 0.4f * cfg.mut.ballMass * ballR * ballR;
        float ii = 1.f / inertia;
        b.invInertiaWorld = M3(V3(ii, 0, 0), V3(0, ii, 0), V3(0, 0, ii));
        V3 ballForce;
        if (ballActive) ballForce += gImp;
        b.extForceImp = ballForce * b.invMass * dt;
        b.extTorqueImp = V3();
        b.dLin = b.dAng = b.push = b.turn = V3();
        b.active = ballActive || ballWoken;
    }
    if (coupled == 0) {
        // the ball is an island of its own: ball-world contacts only, straight from its segment
        if (sb[0].active) {
            if (x.h->nBall > 0) {
                solve_island_one(sb[0], 0, seg_ball(scratch), x.h->nBall);
                a.ball.vel = sb[0].linVel; a.ball.angvel = sb[0].angVel;
                a.ball.pos = sb[0].pos + a.ball.vel * dt;  // integrateTransformNoRot
            } else {
                a.ball.vel = (a.ball.vel + V3()) + sb[0].extForceImp;
                a.ball.angvel = (a.ball.angvel + V3()) + V3();
                a.ball.pos = a.ball.pos + a.ball.vel * dt;
            }
        }
    } else {
        // the ball and the coupled cars: gather in manifold order — ball-world, car-ball (c ascending), then per car:
        // car-world, car-car (c, d > c)
        ContactSet cs; cs.n = 0; cs.overflow = 0;
        auto take = [&](const Contact* src, int n) {
            for (int i = 0; i < n; i++) {
                if (cs.n < kMaxContacts) cs.c[cs.n++] = src[i];
                else cs.overflow++;
            }
        };
        take(seg_ball(scratch), x.h->nBall);
        for (int c = 0; c < P; c++) if ((coupled >> c) & 1u) take(seg_car(scratch, c), x.car[c].nCarBall);
        for (int c = 0; c < P; c++) {
            if (!((coupled >> c) & 1u)) continue;
            take(seg_car_world(scratch, c), car_mesh_count(x.car[c]));
            take(seg_car_plane(scratch, c), car_plane_taken(x.car[c]));
            const Contact* pr = seg_pair(scratch, P);
            for (int i = 0; i < cp.n; i++) if (pr[i].b == 1 + c) take(pr + i, 1);
        }
        for (int c = 0; c < P; c++) {
            if ((coupled >> c) & 1u) solver_body_from_car(sb[1 + c], a.cars[c], x.car[c], k);
            else sb[1 + c].active = 0;  // its own island (or not simulated): contacts that name it are skipped
        }
        solve_island(sb, 1 + P, 0, cs.c, cs.n);
        // write back; the cars integrate themselves in P3
        for (int c = 0; c < P; c++) {
            if (!((coupled >> c) & 1u) || !sb[1 + c].active) continue;
            CarS& car = a.cars[c];
            car.vel = sb[1 + c].linVel; car.angvel = sb[1 + c].angVel;
            car.pos = sb[1 + c].pos; car.rot = sb[1 + c].rot;
        }
        if (sb[0].active) {
            a.ball.vel = sb[0].linVel; a.ball.angvel = sb[0].angVel;
            a.ball.pos = sb[0].pos + a.ball.vel * dt;  // integrateTransformNoRot
        }
    }
    RL_PT(11);
    // Ball::_FinishPhysicsTick (Ball.cpp:112-138)
    V3 ballVelCache;
    for (int c = 0; c < P; c++) ballVelCache += x.car[c].ballVelCache;
    if (!is_zero(ballVelCache)) a.ball.vel += ballVelCache;
    {
        const float maxSpeed = cfg.mut.ballMaxSpeed * UU2BT;
        if (len2(a.ball.vel) > maxSpeed * maxSpeed) a.ball.vel = normalized(a.ball.vel) * maxSpeed;
        if (len2(a.ball.angvel) > C::BALL_MAX_ANG_SPEED * C::BALL_MAX_ANG_SPEED) a.ball.angvel = normalized(a.ball.angvel) * C::BALL_MAX_ANG_SPEED;
    }
    a.ball.updateCounterLo++;
    set_i64(a.tickLo, a.tickHi, tick + 1);
    RL_PT(12);
}

RL_HD inline void tick_p3_car(ArenaS& a, TickX x, const Tables& tb, const CarConsts& k, int c, CarW& w) {
    CarS& car = a.cars[c];
    CarX& o = x.car[c];
    RL_PT(-1);
    if (!o.noResponse) {
        integrate_transform(car.pos, car.rot, car.vel, car.angvel, kTickTime);
        // demolished during this tick: the body still integrates, but Car::_PostTickUpdate returns before it refreshes
        // CarState::rotMat (Car.cpp:133-138), so the rotation every reader sees (GetState -> obs, the next SetState)
        // stays the start-of-tick one until the respawn replaces it
        if (car.isDemoed) car.rot = o.rot;
    }
    w.velCache = o.velCache;
    car_post_tick(car, w);
    uint64_t hit = pads_check_car(a, tb, k, c);
    o.padHitLo = (uint32_t)hit; o.padHitHi = (uint32_t)(hit >> 32);
    RL_PT(14);
}

RL_HD inline void tick_p4_pads(ArenaS& a, TickX x, const SimCfg& cfg) {
    if (cfg.numCars <= 0) return;
    uint64_t hits[kMaxCars];
    for (int c = 0; c < cfg.numCars; c++) hits[c] = ((uint64_t)x.car[c].padHitHi << 32) | x.car[c].padHitLo;
    pads_post_tick(a, cfg, hits);
}

// scratch sizes for one arena (words / contacts)
RL_HDI int tick_scratch_words(int ncars) { return tickx_words(ncars); }

// Serial driver: the same phases, role after role.  Used by the host test build and by single-lane device paths.
// xwords: tickx_words(numCars) uint32 words, scratch: contact_scratch_slots(numCars) contacts.
RL_HD inline void arena_tick(ArenaS& a, const SimCfg& cfg, const MeshSet& ms, const Tables& tb, int firstTickOfStep, uint32_t* xwords, Contact* scratch) {
    const CarConsts k = car_consts(cfg.carPreset);
    const Thresholds thr = contact_thresholds(k, cfg.mut.ballRadius);
    TickX x = make_tickx(xwords);
    const int P = cfg.numCars;
    CarW w[kMaxCars];
    tick_s0_ball(a, x);
    for (int c = 0; c < P; c++) tick_s0_car(a, x, c);
    for (int c = 0; c < P; c++) tick_p1_car(a, x, cfg, ms, k, thr, c, w[c], scratch, firstTickOfStep);
    tick_p1_ball(a, x, cfg, ms, k, thr, scratch);
    bool self[kMaxCars];
    for (int c = 0; c < P; c++) { self[c] = tick_p2_car_self(a, x, cfg, k, c, scratch); if (self[c]) tick_p3_car(a, x, tb, k, c, w[c]); }
    tick_p2_solve(a, x, cfg, k, thr, scratch, firstTickOfStep);
    for (int c = 0; c < P; c++) if (!self[c]) tick_p3_car(a, x, tb, k, c, w[c]);
    tick_p4_pads(a, x, cfg);
#if defined(RL_DEBUG_CONTACTS) && !defined(__CUDA_ARCH__)
    for (int c = 0; c < P; c++) g_dbg_carw[c] = w[c];
#endif
}

// ---- Gym::Step (G/Gym.cpp:68-102) + GameInst::Step auto-reset --------------------------------------
RL_HD inline void gym_step(ArenaS& a, const SimCfg& cfg, const MeshSet& ms, const Tables& tb, const int32_t* actionIdx,
                           float* obsOut, float* rewardOut, uint8_t* doneOut, uint32_t* xwords, Contact* scratch) {
    parse_actions(a, cfg, tb, actionIdx);
    arena_tick(a, cfg, ms, tb, 1, xwords, scratch);
    event_tracker_update(a, cfg);
    snapshot_update(a, cfg);
    build_obs(a, cfg, tb, obsOut);
    bool done = compute_done(a, cfg);
    compute_rewards(a, cfg, rewardOut);
    *doneOut = done ? 1 : 0;
    for (int t = 1; t < cfg.tickSkip; t++) arena_tick(a, cfg, ms, tb, 0, xwords, scratch);
    if (done) {
        gym_reset(a, cfg);
        build_obs(a, cfg, tb, obsOut);
    }
}

}  // namespace rl
