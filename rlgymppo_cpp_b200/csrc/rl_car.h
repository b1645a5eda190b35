// rl_car.h — Car::_PreTickUpdate / _PostTickUpdate / _FinishPhysicsTick and btVehicleRL as device code.
//
// Reference: R/Sim/Car/Car.cpp:58-834 (control -> force model) and
// R/Sim/btVehicleRL/btVehicleRL.cpp:118-402 (4-wheel suspension raycast + friction).
// Order of force/impulse application inside a tick follows SURVEY.md appendix A2.
#pragma once
#include "rl_mesh.h"

namespace rl {

struct WheelW {
    V3 hardPoint, contactPoint, contactNormal, axle, impulse;
    int32_t inContact, inContactWorld, ground;  // ground: -2 none, -1 static, >=0 dynamic body
    float suspLen, suspRelVel, clippedInv, suspForce;
};

struct CarW {
    V3 force, torque, velCache;
    M3 invInertiaWorld;
    WheelW w[4];
    MeshCands cands;  // this tick's mesh triangles near the car (wheel rays + hitbox), see rl_mesh.h
    RayHit meshHit[4];  // the four wheel rays against the mesh candidates (wheel_mesh_rays), the start of wheel_ray
    // hitbox narrowphase pre-filter from the same pass: candidates whose leaf box overlaps the hitbox AABB, and the first
    // candidate of every mesh (haveMask = 0: not computed, the narrowphase takes its serial path)
    uint32_t candMask, candGroupStart;
    int32_t haveMask;
    // btContactSolverInfo::m_timeStep as this tick's vehicle update sees it: the tick time — except during the very first tick
    // of an arena's life (tick count 0; the count only ever grows, Arena.cpp:810), when the field still holds its constructor
    // default 1/60: it is set inside stepSimulation, after Car::_PreTickUpdate.  The wheels' extra push-back reads it
    // (btVehicleRL.cpp:183-199 -> resolveSingleCollision).
    float solverDt;
};

// ---- per-arena exchange between the roles of a tick (ball role + one role per car) --------------------------------
// On the device this lives in shared memory next to the arena state (engine.cu k_roles); the host test build uses a
// plain buffer.  The snapshot members are written at the start of a tick and are what OTHER roles read while the
// owner updates its body (wheel rays against other cars / the ball, car-ball tests), so the result does not depend on
// the order in which the roles run.
struct CarX {
    V3 pos, vel, angvel;  // start-of-tick snapshot
    M3 rot;
    int32_t demoed;
    V3 force, torque;     // accumulated by Car::_PreTickUpdate (+ gravity), consumed by the solver
    V3 cmn, cmx;          // hitbox AABB after the pre-tick
    V3 ballVelCache;      // this car's contribution to Ball::_velocityImpulseCache
    V3 velCache;          // Car::_velocityImpulseCache (bumps), written by the pair phase
    int32_t noResponse;   // demoed when the tick started
    int32_t nCarBall, nCarPlane;  // contacts written to this car's scratch segments (car-ball slot, plane staging slots)
    // P3 -> P4: the boost pads this car overlaps.  Between P1 and P3 the low word is free and carries the number of hitbox-mesh
    // contacts in the car's world slots (car_mesh_count), written by whichever role ran that narrowphase.
    uint32_t padHitLo, padHitHi;
};
RL_HDI int car_mesh_count(const CarX& o) { return (int)o.padHitLo; }
RL_HDI void car_set_mesh_count(CarX& o, int n) { o.padHitLo = (uint32_t)n; }
struct TickXHdr {
    V3 ballPos, ballVel, ballAngvel;  // start-of-tick snapshot (vel undamped)
    int32_t nBall, nPair, ballActive;
};
constexpr int kTickXHdrWords = sizeof(TickXHdr) / 4;
constexpr int kCarXWords = sizeof(CarX) / 4;
RL_HDI int tickx_words(int ncars) { return kTickXHdrWords + ncars * kCarXWords; }
struct TickX {
    TickXHdr* h;
    CarX* car;
};
RL_HDI TickX make_tickx(uint32_t* words) {
    TickX x;
    x.h = reinterpret_cast<TickXHdr*>(words);
    x.car = reinterpret_cast<CarX*>(words + kTickXHdrWords);
    return x;
}

// constants derived the way Car::_BulletSetup derives them (Car.cpp:195-283)
struct CarConsts {
    V3 halfExt;      // half extents WITH margin (btBoxShape::getHalfExtentsWithMargin), Bullet units
    V3 coreHalf;     // m_implicitShapeDimensions = requested half extents - 0.04
    float boxMargin; // collision margin after setSafeMargin: 0.1 * smallest half extent (< 0.04 for every preset)
    V3 hitboxOffset; // child transform origin
    V3 invInertiaLocal;
    float invMass;
    V3 wheelConn[4];
    float wheelRadius[4], wheelRest[4], wheelForceScale[4];
    float suspTravel;  // m_maxSuspensionTravelCm / 100
};

RL_HDI CarConsts car_consts(int preset = 0) {
    CarConsts k;
    const C::CarPreset cp = C::car_preset(preset);
    // btBoxShape ctor (B/BulletCollision/CollisionShapes/btBoxShape.cpp:18-28): implicit dims = half - 0.04, then
    // setSafeMargin lowers ONLY the margin to 0.1 * min half extent (setMargin is not virtual in this fork), so
    // the effective box is 0.04 - margin smaller than the configured hitbox on every side.
    V3 req((cp.hitbox[0] * UU2BT) / 2, (cp.hitbox[1] * UU2BT) / 2, (cp.hitbox[2] * UU2BT) / 2);
    k.coreHalf = req - V3(C::BOX_MARGIN, C::BOX_MARGIN, C::BOX_MARGIN);
    float minDim = fminf_(fminf_(req.x, req.y), req.z);
    float safe = 0.1f * minDim;
    k.boxMargin = safe < C::BOX_MARGIN ? safe : C::BOX_MARGIN;
    k.halfExt = k.coreHalf + V3(k.boxMargin, k.boxMargin, k.boxMargin);
    k.hitboxOffset = V3(cp.hitboxOff[0] * UU2BT, cp.hitboxOff[1] * UU2BT, cp.hitboxOff[2] * UU2BT);
    // btBoxShape::calculateLocalInertia
    float lx = 2.f * k.halfExt.x, ly = 2.f * k.halfExt.y, lz = 2.f * k.halfExt.z;
    V3 inertia(C::CAR_MASS / 12.f * (ly * ly + lz * lz), C::CAR_MASS / 12.f * (lx * lx + lz * lz), C::CAR_MASS / 12.f * (lx * lx + ly * ly));
    k.invInertiaLocal = V3(1.f / inertia.x, 1.f / inertia.y, 1.f / inertia.z);
    k.invMass = 1.f / C::CAR_MASS;
    for (int i = 0; i < 4; i++) {
        bool front = i < 2, left = (i % 2) != 0;
        V3 off = front ? V3(cp.wheelF[0], cp.wheelF[1], cp.wheelF[2]) : V3(cp.wheelB[0], cp.wheelB[1], cp.wheelB[2]);
        if (left) off.y *= -1.f;
        k.wheelConn[i] = V3(off.x * UU2BT, off.y * UU2BT, off.z * UU2BT);
        k.wheelRadius[i] = (front ? cp.wheelRFront : cp.wheelRBack) * UU2BT;
        float rest = front ? cp.susRestFront : cp.susRestBack;
        rest -= C::MAX_SUSPENSION_TRAVEL;
        k.wheelRest[i] = rest * UU2BT;
        k.wheelForceScale[i] = front ? C::SUSPENSION_FORCE_SCALE_FRONT : C::SUSPENSION_FORCE_SCALE_BACK;
    }
    k.suspTravel = ((C::MAX_SUSPENSION_TRAVEL * UU2BT) * 100) / 100;
    return k;
}

RL_HDI V3 vel_at(V3 lin, V3 ang, V3 rel) { return lin + cross(ang, rel); }

// btRigidBody::applyImpulse
RL_HDI void apply_impulse(CarS& c, const CarW& w, float invMass, V3 impulse, V3 rel) {
    c.vel += impulse * invMass;
    c.angvel += w.invInertiaWorld * cross(rel, impulse);
}

// btRigidBody::computeImpulseDenominator
RL_HDI float impulse_denom(V3 bodyPos, const M3& invInertiaWorld, float invMass, V3 pos, V3 normal) {
    V3 r0 = pos - bodyPos;
    V3 c0 = cross(r0, normal);
    V3 vec = cross(tmul(c0, invInertiaWorld), r0);
    return invMass + dot(normal, vec);
}

// minimal view of "the other dynamic bodies" for wheel rays / wheel-on-body friction
struct DynView {
    V3 pos, vel, angvel;
    M3 rot;
    M3 invInertiaWorld;
    V3 invInertiaLocal;
    float invMass;
    int32_t responds;  // hasContactResponse (false for demoed cars)
};

RL_HD inline DynView dyn_view(const TickX& x, int body, const CarConsts& k, const Mut& mu) {
    DynView d;
    if (body == 0) {
        d.pos = x.h->ballPos; d.vel = x.h->ballVel; d.angvel = x.h->ballAngvel; d.rot = M3::identity();
        float r = mu.ballRadius * UU2BT;
        float inertia = 0.4f * mu.ballMass * r * r;
        d.invInertiaLocal = V3(1.f / inertia, 1.f / inertia, 1.f / inertia);
        d.invMass = 1.f / mu.ballMass; d.responds = 1;
    } else {
        const CarX& c = x.car[body - 1];
        d.pos = c.pos; d.vel = c.vel; d.angvel = c.angvel; d.rot = c.rot;
        d.invInertiaLocal = k.invInertiaLocal; d.invMass = k.invMass; d.responds = !c.demoed;
    }
    d.invInertiaWorld = world_inertia(d.rot, d.invInertiaLocal);
    return d;
}

// btCollisionWorld::rayTest through btRSBroadphase::rayTest for one wheel ray
// the wheel rays of a car (btVehicleRL.cpp:118-216 rayCast): wheel i from its hard point down the suspension
RL_HDI void wheel_ray_segments(const CarS& c, const CarConsts& k, V3* from, V3* to) {
    V3 wheelDir = c.rot * V3(0, 0, -1);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float rayLen = k.wheelRest[i] + k.suspTravel + k.wheelRadius[i] - C::SUSPENSION_SUBTRACTION;
        from[i] = c.pos + c.rot * k.wheelConn[i];
        to[i] = from[i] + wheelDir * rayLen;
    }
}
RL_HDI void wheel_mesh_rays_init(CarW& w) {
    for (int i = 0; i < 4; i++) { w.meshHit[i].frac = 1.0f; w.meshHit[i].body = -2; w.meshHit[i].normal = V3(0, 0, 1); }
    w.candMask = 0; w.candGroupStart = 0; w.haveMask = 0;
}
// Serial form of the candidate pass: the mesh part of the four wheel rays (triangle-major over the candidate list, or the
// direct BVH walks when the list overflowed).  The role kernel does the same per (car, candidate) pair across the warp
// (engine.cu cands_pass_warp).
RL_HD inline void wheel_mesh_rays(const CarS& c, const CarConsts& k, const MeshSet& ms, CarW& w) {
    wheel_mesh_rays_init(w);
    if (c.isDemoed) return;  // Car::_PreTickUpdate returns before the vehicle update
    V3 from[4], to[4];
    wheel_ray_segments(c, k, from, to);
    if (w.cands.n >= 0) {
        for (int j = 0; j < w.cands.n; j++) {
            const BvhNode& nd = ms.nodes[w.cands.node[j] & 0xffffff];
            const Tri& t = ms.tris[nd.tri];
            TriRays r;
            ray_tri4(from, to, t.v0, t.v1, t.v2, r);
            ray_tri4_apply(r.d, r.nn, r.neg, w.meshHit);
        }
    } else {
        for (int i = 0; i < 4; i++) ray_meshes(from[i], to[i], ms, w.meshHit[i]);
    }
}

RL_HD inline RayHit wheel_ray(const TickX& x, const SimCfg& cfg, const MeshSet& ms, const CarConsts& k, const RayHit& meshHit, int self, V3 from, V3 to) {
    RayHit hit = meshHit;  // static meshes first (wheel_mesh_rays), then planes, ball, other cars
    for (int p = 0; p < 4; p++) ray_plane(from, to, world_plane(p), hit);
    ray_sphere(from, to, x.h->ballPos, cfg.mut.ballRadius * UU2BT, 0, hit);
    for (int c = 0; c < cfg.numCars; c++) {
        if (c == self) continue;
        const CarX& o = x.car[c];
        V3 center = o.pos + o.rot * k.hitboxOffset;
        ray_obb(from, to, center, o.rot, k.halfExt, 1 + c, hit);
    }
    if (hit.body >= 1 && x.car[hit.body - 1].demoed) hit.body = -2;  // btDefaultVehicleRaycaster.cpp:40-51
    return hit;
}

// ---- btVehicleRL::updateVehicleFirst ---------------------------------------------------------
RL_HD inline void vehicle_first(CarS& c, const TickX& x, const SimCfg& cfg, const MeshSet& ms, const CarConsts& k, int ci, CarW& w) {
    V3 carUp = c.rot.col(2);
    // updateWheelTransformsWS + updateWheelTransform: only the steered axle (basis column 1) is consumed later.  The two
    // front wheels share one steering rotation; a zero angle (the rear wheels) gives the exact identity rotation
    // (quat_axis_angle(up, 0) = (0, 0, 0, 1)), so its axle is -wheelAxle itself.
    const V3 wheelDir = c.rot * V3(0, 0, -1);
    const V3 wheelAxle = c.rot * V3(0, -1, 0);
    const V3 rearAxle = -wheelAxle;
    const V3 frontAxle = (c.wheelSteer == 0.f) ? rearAxle : quat_to_mat(quat_axis_angle(-wheelDir, c.wheelSteer)) * (-wheelAxle);
    RL_WHEEL_LOOP
    for (int i = 0; i < 4; i++) {
        WheelW& wh = w.w[i];
        wh.hardPoint = c.pos + c.rot * k.wheelConn[i];
        wh.axle = (i < 2) ? frontAxle : rearAxle;

        // rayCast (btVehicleRL.cpp:118-216)
        float rayLen = k.wheelRest[i] + k.suspTravel + k.wheelRadius[i] - C::SUSPENSION_SUBTRACTION;
        V3 source = wh.hardPoint;
        V3 target = source + wheelDir * rayLen;
        wh.contactPoint = target;
        wh.ground = -2; wh.inContact = 0; wh.inContactWorld = 0;
        RayHit hit = wheel_ray(x, cfg, ms, k, w.meshHit[i], ci, source, target);
        if (hit.body != -2) {
            wh.contactPoint = source + (target - source) * hit.frac;
            wh.contactNormal = hit.normal;
            wh.inContact = 1;
            wh.inContactWorld = hit.body == -1;
            wh.ground = hit.body;
            float traceLen = dot(wh.hardPoint - wh.contactPoint, carUp);
            wh.suspLen = clampf(traceLen - k.wheelRadius[i], k.wheelRest[i] - k.suspTravel, k.wheelRest[i] + k.suspTravel);
            float denom = dot(wh.contactNormal, carUp);
            V3 relpos = wh.contactPoint - c.pos;
            V3 velAt = vel_at(c.vel, c.angvel, relpos);
            float projVel = dot(wh.contactNormal, velAt);
            if (denom > 0.1f) {
                float inv = 1.f / denom;
                wh.suspRelVel = projVel * inv;
                wh.clippedInv = inv;
            } else {
                wh.suspRelVel = 0.f;
                wh.clippedInv = 10.f;
            }
            if (hit.body == -1) {
                float thresh = (k.wheelRest[i] + k.wheelRadius[i]) - C::SUSPENSION_SUBTRACTION;
                if (traceLen < thresh) {
                    // resolveSingleCollision(chassis, static, hitPoint, hitNormal, solverInfo, delta, false)
                    // (B/BulletDynamics/ConstraintSolver/btContactConstraint.cpp:60-106; m_erp = 0.2)
                    float delta = traceLen - thresh;
                    float rel_vel = dot(hit.normal, velAt);
                    float positionalError = 0.2f * -delta / w.solverDt;
                    float velocityError = -(1.0f + 0.f) * rel_vel;
                    float denom0 = impulse_denom(c.pos, w.invInertiaWorld, k.invMass, wh.contactPoint, hit.normal);
                    float jacDiagABInv = 1.f / (denom0 + 0.f);
                    float normalImpulse = positionalError * jacDiagABInv + velocityError * jacDiagABInv;
                    normalImpulse = 0.f > normalImpulse ? 0.f : normalImpulse;
                    c.wheelPush[i] = normalImpulse / 4;
                }
            }
        } else {
            wh.suspLen = k.wheelRest[i] + k.suspTravel;
            wh.suspRelVel = 0.f;
            wh.contactNormal = -wheelDir;
            wh.clippedInv = 1.f;
            c.wheelPush[i] = 0.f;
        }
    }

    // calcFrictionImpulses (btVehicleRL.cpp:313-388) — consumes LAST tick's engine/brake/friction values
    float frictionScale = C::CAR_MASS / 3;
    RL_WHEEL_LOOP
    for (int i = 0; i < 4; i++) {
        WheelW& wh = w.w[i];
        if (wh.ground == -2) { wh.impulse = V3(); continue; }
        V3 axleDir = wh.axle;
        V3 surf = wh.contactNormal;
        float proj = dot(axleDir, surf);
        axleDir -= surf * proj;
        axleDir = safe_normalized(axleDir);
        V3 forwardDir = safe_normalized(cross(surf, axleDir));

        // resolveSingleBilateral (btContactConstraint.cpp:108-157)
        DynView g;
        bool dynGround = wh.ground >= 0;
        if (dynGround) g = dyn_view(x, wh.ground, k, cfg.mut);
        V3 rel1 = wh.contactPoint - c.pos;
        V3 rel2 = dynGround ? (wh.contactPoint - g.pos) : V3();
        V3 vel1 = vel_at(c.vel, c.angvel, rel1);
        V3 vel2 = dynGround ? vel_at(g.vel, g.angvel, rel2) : V3();
        V3 vel = vel1 - vel2;
        V3 aJ = tmul(cross(rel1, axleDir), c.rot);  // world2A * (rel_pos1 x axis), world2A = basis^T
        V3 minvJt0 = k.invInertiaLocal * aJ;
        float diag = k.invMass + dot(minvJt0, aJ);
        if (dynGround) {
            V3 bJ = tmul(cross(rel2, -axleDir), g.rot);
            V3 minvJt1 = g.invInertiaLocal * bJ;
            diag += g.invMass + dot(minvJt1, bJ);
        }
        float sideImpulse = -0.2f * dot(axleDir, vel) * (1.f / diag);

        float rollingFriction;
        if (c.wheelEngine == 0.f) {
            if (c.wheelBrake != 0.f) {
                V3 carRel = wh.contactPoint - c.pos;
                V3 v1 = vel_at(c.vel, c.angvel, carRel);
                V3 v2 = dynGround ? vel_at(g.vel, g.angvel, carRel) : V3();  // (sic) reference uses the car-relative point
                float relVel = dot(v1 - v2, forwardDir);
                const float MAGIC = 113.73963f;
                rollingFriction = clampf(-relVel * MAGIC, -c.wheelBrake, c.wheelBrake);
            } else {
                rollingFriction = 0.f;
            }
        } else {
            rollingFriction = -c.wheelEngine / frictionScale;
        }
        V3 total = (forwardDir * rollingFriction * c.wheelLong[i]) + (axleDir * sideImpulse * c.wheelLat[i]);
        wh.impulse = total * frictionScale;
    }
}

RL_HDI V3 up_from_wheel_contacts(const CarS& c, const CarW& w) {
    V3 sum(0, 0, 0);
    for (int i = 0; i < 4; i++) if (w.w[i].inContact) sum += w.w[i].contactNormal;
    if (is_zero(sum)) return c.rot.col(2);
    return safe_normalized(sum);
}

// ---- Car::_UpdateWheels (Car.cpp:330-475) ----------------------------------------------------
RL_HD inline void update_wheels(CarS& c, CarW& w, int numWheelsInContact, float forwardSpeedUU) {
    const float dt = kTickTime;
    float absFwd = fabsf(forwardSpeedUU);
    bool wheelsWorld = false;
    for (int i = 0; i < 4; i++) wheelsWorld |= w.w[i].inContactWorld != 0;
    if (c.controls.handbrake) c.handbrakeVal += C::POWERSLIDE_RISE_RATE * dt;
    else c.handbrakeVal -= C::POWERSLIDE_FALL_RATE * dt;
    c.handbrakeVal = clampf(c.handbrakeVal, 0.f, 1.f);

    float realThrottle = c.controls.throttle;
    float realBrake = 0;
    if (c.controls.boost && c.boost > 0) realThrottle = 1;
    {
        const float tx[3] = {0, 1400, 1410}, ty[3] = {1.0f, 0.1f, 0.0f};
        float driveSpeedScale = curve(tx, ty, absFwd);
        float engineThrottle = realThrottle;
        if (!c.controls.handbrake) {
            float absThrottle = fabsf(realThrottle);
            if (absThrottle >= C::THROTTLE_DEADZONE) {
                if (absFwd > C::STOPPING_FORWARD_VEL && sgn(realThrottle) != sgn(forwardSpeedUU)) {
                    realBrake = 1;
                    if (absFwd > C::BRAKING_NO_THROTTLE_SPEED_THRESH) engineThrottle = 0;
                }
            } else {
                engineThrottle = 0;
                bool fullStop = absFwd < C::STOPPING_FORWARD_VEL;
                realBrake = fullStop ? 1 : C::COASTING_BRAKE_FACTOR;
            }
        }
        if (numWheelsInContact < 3) driveSpeedScale /= 4;
        c.wheelEngine = engineThrottle * (C::THROTTLE_TORQUE_AMOUNT * UU2BT) * driveSpeedScale;
        c.wheelBrake = realBrake * (C::BRAKE_TORQUE_AMOUNT * UU2BT);
    }
    {
        const float sx[6] = {0, 500, 1000, 1500, 1750, 3000};
        const float sy[6] = {0.53356f, 0.31930f, 0.18203f, 0.10570f, 0.08507f, 0.03454f};
        float steerAngle = curve(sx, sy, absFwd);
        if (c.handbrakeVal != 0.f) {
            const float px[2] = {0, 2500}, py[2] = {0.39235f, 0.12610f};
            steerAngle += (curve(px, py, absFwd) - steerAngle) * c.handbrakeVal;
        }
        steerAngle *= c.controls.steer;
        c.wheelSteer = steerAngle;
    }
    RL_WHEEL_LOOP
    for (int i = 0; i < 4; i++) {
        WheelW& wh = w.w[i];
        if (wh.ground == -2) continue;
        V3 latDir = wh.axle;
        V3 longDir = cross(latDir, wh.contactNormal);
        float frictionCurveInput = 0;
        V3 wheelDelta = wh.hardPoint - c.pos;
        V3 crossVec = (cross(c.angvel, wheelDelta) + c.vel) * BT2UU;
        float baseFriction = fabsf(dot(crossVec, latDir));
        if (baseFriction > 5) frictionCurveInput = baseFriction / (fabsf(dot(crossVec, longDir)) + baseFriction);
        const float lx[2] = {0, 1}, ly[2] = {1.0f, 0.2f};
        float latFriction = curve(lx, ly, frictionCurveInput);
        float longFriction = 1;  // LONG_FRICTION_CURVE is empty -> default output 1
        if (c.handbrakeVal != 0.f) {
            float hb = c.handbrakeVal;
            const float hlx[2] = {0, 1}, hly[2] = {0.5f, 0.9f};
            latFriction *= (0.1f - 1) * hb + 1;  // HANDBRAKE_LAT_FRICTION_FACTOR_CURVE is the constant 0.1
            longFriction *= (curve(hlx, hly, frictionCurveInput) - 1) * hb + 1;
        } else {
            longFriction = 1;
        }
        bool sticky = realThrottle != 0;
        if (!sticky) {
            const float nx[3] = {0, 0.7075f, 1}, ny[3] = {0.1f, 0.5f, 1.0f};
            float s = curve(nx, ny, wh.contactNormal.z);
            latFriction *= s; longFriction *= s;
        }
        c.wheelLat[i] = latFriction;
        c.wheelLong[i] = longFriction;
    }
    if (wheelsWorld) {
        V3 upDir = up_from_wheel_contacts(c, w);
        bool fullStick = (realThrottle != 0) || (absFwd > C::STOPPING_FORWARD_VEL);
        float stickyScale = 0.5f;
        if (fullStick) stickyScale += 1 - fabsf(upDir.z);
        w.force += upDir * stickyScale * (C::GRAVITY_Z * UU2BT) * C::CAR_MASS;
    }
}

RL_HDI M3 inertia_world(const CarS& c, const CarConsts& k) {
    V3 I(1.f / k.invInertiaLocal.x, 1.f / k.invInertiaLocal.y, 1.f / k.invInertiaLocal.z);
    return world_inertia(c.rot, I);
}

// ---- Car::_UpdateAirTorque (Car.cpp:556-641) -------------------------------------------------
RL_HD inline void update_air_torque(CarS& c, CarW& w, const CarConsts& k, bool updateAirControl) {
    V3 dirPitch = -c.rot.col(1), dirYaw = c.rot.col(2), dirRoll = -c.rot.col(0);
    bool doAirControl = false;
    if (c.isFlipping) c.isFlipping = c.hasFlipped && c.flipTime < C::FLIP_TORQUE_TIME;
    M3 Iw = inertia_world(c, k);
    if (c.isFlipping) {
        V3 rel = c.flipRelTorque;
        if (!is_zero(c.flipRelTorque)) {
            float pitchScale = 1;
            if (rel.y != 0 && c.controls.pitch != 0) {
                if (sgn(rel.y) == sgn(c.controls.pitch)) {
                    pitchScale = 1 - fminf_(fabsf(c.controls.pitch), 1.f);
                    doAirControl = true;
                }
            }
            rel.y *= pitchScale;
            V3 dodgeTorque = rel * V3(C::FLIP_TORQUE_X, C::FLIP_TORQUE_Y, 0);
            w.torque += (Iw * c.rot) * dodgeTorque;
        } else {
            doAirControl = true;
        }
    } else {
        doAirControl = true;
    }
    doAirControl &= !c.isAutoFlipping;
    doAirControl &= updateAirControl;
    if (doAirControl) {
        float pitchTorqueScale = 1;
        V3 torque;
        if (c.controls.pitch != 0 || c.controls.yaw != 0 || c.controls.roll != 0) {
            if (c.isFlipping) pitchTorqueScale = 0;
            else if (c.hasFlipped) { if (c.flipTime < C::FLIP_TORQUE_TIME + C::FLIP_PITCHLOCK_EXTRA_TIME) pitchTorqueScale = 0; }
            torque = (dirPitch * c.controls.pitch * pitchTorqueScale * C::AIR_TORQUE_P) + (dirYaw * c.controls.yaw * C::AIR_TORQUE_Y) +
                     (dirRoll * c.controls.roll * C::AIR_TORQUE_R);
        } else {
            torque = V3(0, 0, 0);
        }
        V3 av = c.angvel;
        float dampPitch = dot(dirPitch, av) * C::AIR_DAMP_P * (1 - fabsf(c.controls.pitch * pitchTorqueScale));
        float dampYaw = dot(dirYaw, av) * C::AIR_DAMP_Y * (1 - fabsf(c.controls.yaw));
        float dampRoll = dot(dirRoll, av) * C::AIR_DAMP_R;
        V3 damping = (dirYaw * dampYaw) + (dirPitch * dampPitch) + (dirRoll * dampRoll);
        w.torque += (Iw * (torque - damping)) * C::CAR_TORQUE_SCALE;
    }
    if (c.controls.throttle != 0) w.force += c.rot.col(0) * c.controls.throttle * C::THROTTLE_AIR_ACCEL * UU2BT * C::CAR_MASS;
}

// ---- Car::_UpdateJump (Car.cpp:507-554) -------------------------------------------------------
RL_HD inline void update_jump(CarS& c, CarW& w, const CarConsts& k, const SimCfg& cfg, bool jumpPressed) {
    const float dt = kTickTime;
    if (c.isOnGround && !c.isJumping) {
        if (c.hasJumped && c.jumpTime < C::JUMP_MIN_TIME + C::JUMP_RESET_TIME_PAD) {
        } else {
            c.hasJumped = 0; c.jumpTime = 0;
        }
    }
    if (c.isJumping) {
        if (c.jumpTime < C::JUMP_MIN_TIME || (c.controls.jump && c.jumpTime < C::JUMP_MAX_TIME)) c.isJumping = 1;
        else c.isJumping = 0;
    } else if (c.isOnGround && jumpPressed) {
        c.isJumping = 1; c.jumpTime = 0;
        V3 imp = c.rot.col(2) * cfg.mut.jumpImmediateForce * UU2BT * C::CAR_MASS;
        c.vel += imp * k.invMass;
    }
    if (c.isJumping) {
        c.hasJumped = 1;
        V3 f = c.rot.col(2) * cfg.mut.jumpAccel;
        if (c.jumpTime < C::JUMP_MIN_TIME) f *= 0.62f;
        w.force += f * UU2BT * C::CAR_MASS;
    }
    if (c.isJumping || c.hasJumped) c.jumpTime += dt;
}

// btMatrix3x3::getEulerYPR -> Angle::FromRotMat (MathTypes.cpp:62-71); only roll is consumed
RL_HDI float rotmat_roll(const M3& rot) {
    // bulletMat[i][j] = rotMat[j][i] where rotMat rows are forward/right/up vectors == our basis
    float pitch = rl_asin(clampf(-rot.r[2].x, -1.f, 1.f));
    float roll = rl_atan2(rot.r[2].y, rot.r[2].z);
    if (fabsf(pitch) == kHalfPi) { if (roll > 0) roll -= kPi; else roll += kPi; }
    return roll * -1.f;
}

// ---- Car::_UpdateAutoFlip (Car.cpp:763-797) ---------------------------------------------------
RL_HD inline void update_auto_flip(CarS& c, const CarConsts& k, bool jumpPressed) {
    const float dt = kTickTime;
    if (jumpPressed && c.worldContactHas && c.worldContactNormal.z > C::CAR_AUTOFLIP_NORMZ_THRESH) {
        float roll = rotmat_roll(c.rot);
        float absRoll = fabsf(roll);
        if (absRoll > C::CAR_AUTOFLIP_ROLL_THRESH) {
            c.autoFlipTimer = C::CAR_AUTOFLIP_TIME * (absRoll / kPi);
            c.autoFlipTorqueScale = (roll > 0) ? 1.f : -1.f;
            c.isAutoFlipping = 1;
            V3 imp = -c.rot.col(2) * C::CAR_AUTOFLIP_IMPULSE * UU2BT * C::CAR_MASS;
            c.vel += imp * k.invMass;
        }
    }
    if (c.isAutoFlipping) {
        if (c.autoFlipTimer <= 0) { c.isAutoFlipping = 0; c.autoFlipTimer = 0; }
        else {
            c.angvel += c.rot.col(0) * C::CAR_AUTOFLIP_TORQUE * c.autoFlipTorqueScale * dt;
            c.autoFlipTimer -= dt;
        }
    }
}

// ---- Car::_UpdateDoubleJumpOrFlip (Car.cpp:643-761) -------------------------------------------
RL_HD inline void update_double_jump_or_flip(CarS& c, const CarConsts& k, const SimCfg& cfg, bool jumpPressed, float forwardSpeedUU) {
    const float dt = kTickTime;
    if (c.isOnGround) {
        c.hasDoubleJumped = 0; c.hasFlipped = 0; c.airTime = 0; c.airTimeSinceJump = 0; c.flipTime = 0;
    } else {
        c.airTime += dt;
        if (c.hasJumped && !c.isJumping) c.airTimeSinceJump += dt; else c.airTimeSinceJump = 0;
        if (jumpPressed && c.airTimeSinceJump < C::DOUBLEJUMP_MAX_DELAY) {
            float inputMagnitude = fabsf(c.controls.yaw) + fabsf(c.controls.pitch) + fabsf(c.controls.roll);
            bool isFlipInput = inputMagnitude >= C::DODGE_DEADZONE;
            bool canUse = (!c.hasDoubleJumped && !c.hasFlipped) || (isFlipInput ? cfg.mut.unlimitedFlips : cfg.mut.unlimitedDoubleJumps);  // Car.cpp:665-671
            if (c.isAutoFlipping) canUse = false;
            if (canUse) {
                if (isFlipInput) {
                    c.flipTime = 0; c.hasFlipped = 1; c.isFlipping = 1;
                    float forwardSpeedRatio = fabsf(forwardSpeedUU) / C::CAR_MAX_SPEED;
                    V3 dodgeDir(-c.controls.pitch, c.controls.yaw + c.controls.roll, 0);
                    if (fabsf(c.controls.yaw + c.controls.roll) < 0.1f && fabsf(c.controls.pitch) < 0.1f) dodgeDir = V3(0, 0, 0);
                    else dodgeDir = safe_normalized(dodgeDir);
                    c.flipRelTorque = V3(-dodgeDir.y, dodgeDir.x, 0);
                    if (fabsf(dodgeDir.x) < 0.1f) dodgeDir.x = 0;
                    if (fabsf(dodgeDir.y) < 0.1f) dodgeDir.y = 0;
                    if (!(len2(dodgeDir) < kEps * kEps)) {  // !fuzzyZero()
                        bool backwards;
                        if (fabsf(forwardSpeedUU) < 100.0f) backwards = dodgeDir.x < 0.0f;
                        else backwards = (dodgeDir.x >= 0.0f) != (forwardSpeedUU >= 0.0f);
                        V3 v = dodgeDir * C::FLIP_INITIAL_VEL_SCALE;
                        float maxScaleX = backwards ? C::FLIP_BACKWARD_IMPULSE_MAX_SPEED_SCALE : C::FLIP_FORWARD_IMPULSE_MAX_SPEED_SCALE;
                        v.x *= ((maxScaleX - 1) * forwardSpeedRatio) + 1.f;
                        v.y *= ((C::FLIP_SIDE_IMPULSE_MAX_SPEED_SCALE - 1) * forwardSpeedRatio) + 1.f;
                        if (backwards) v.x *= C::FLIP_BACKWARD_IMPULSE_SCALE_X;
                        V3 fwd = c.rot.col(0);
                        float ang = rl_atan2(fwd.y, fwd.x);
                        float ca = rl_cos(ang), sa = rl_sin(ang);
                        V3 xDir(ca, -sa, 0.f), yDir(sa, ca, 0.f);
                        V3 dv(dot(v, xDir), dot(v, yDir), 0.f);
                        c.vel += (dv * UU2BT * C::CAR_MASS) * k.invMass;
                    }
                } else {
                    V3 imp = c.rot.col(2) * C::JUMP_IMMEDIATE_FORCE * UU2BT * C::CAR_MASS;
                    c.vel += imp * k.invMass;
                    c.hasDoubleJumped = 1;
                }
            }
        }
    }
    if (c.isFlipping) {
        c.flipTime += dt;
        if (c.flipTime <= C::FLIP_TORQUE_TIME) {
            if (c.flipTime >= C::FLIP_Z_DAMP_START && (c.vel.z < 0 || c.flipTime < C::FLIP_Z_DAMP_END)) c.vel.z *= cfg.flipZDampFactor;
        }
    } else if (c.hasFlipped) {
        c.flipTime += dt;
    }
}

// ---- Car::_UpdateAutoRoll (Car.cpp:799-833) ---------------------------------------------------
RL_HD inline void update_auto_roll(CarS& c, CarW& w, const CarConsts& k, int numWheelsInContact) {
    V3 groundUp = numWheelsInContact > 0 ? up_from_wheel_contacts(c, w) : c.worldContactNormal;
    V3 groundDown = -groundUp;
    V3 fwd = c.rot.col(0), right = c.rot.col(1);
    V3 crossRight = cross(groundUp, fwd), crossFwd = cross(groundDown, crossRight);
    float rightFactor = 1 - clampf(dot(right, crossRight), 0.f, 1.f);
    float fwdFactor = 1 - clampf(dot(fwd, crossFwd), 0.f, 1.f);
    V3 torqueDirRight = fwd * (dot(right, groundUp) >= 0 ? -1.f : 1.f);
    V3 torqueDirFwd = right * (dot(fwd, groundUp) >= 0 ? 1.f : -1.f);
    V3 tRight = torqueDirRight * rightFactor, tFwd = torqueDirFwd * fwdFactor;
    w.force += groundDown * C::CAR_AUTOROLL_FORCE * UU2BT * C::CAR_MASS;
    w.torque += (inertia_world(c, k) * (tFwd + tRight)) * C::CAR_AUTOROLL_TORQUE;
}

// ---- btVehicleRL::updateVehicleSecond (btVehicleRL.cpp:237-311,390-402) -----------------------
RL_HD inline void vehicle_second(CarS& c, CarW& w, const CarConsts& k) {
    const float dt = kTickTime;
    // the chassis velocities and the inertia tensor stay in registers across the eight applyImpulse calls (same
    // operations in the same order as btRigidBody::applyImpulse on the body)
    const V3 pos = c.pos;
    V3 vel = c.vel, angvel = c.angvel;
    const M3 iiw = w.invInertiaWorld;
    RL_WHEEL_LOOP
    for (int i = 0; i < 4; i++) {
        WheelW& wh = w.w[i];
        float suspForce = 0;
        if (wh.inContact) {
            float force = (k.wheelRest[i] - wh.suspLen) * C::SUSPENSION_STIFFNESS * wh.clippedInv;
            float damp = (wh.suspRelVel < 0) ? C::WHEELS_DAMPING_COMPRESSION : C::WHEELS_DAMPING_RELAXATION;
            suspForce = force - (damp * wh.suspRelVel);
            suspForce *= k.wheelForceScale[i];
            if (suspForce < 0) suspForce = 0;
        }
        wh.suspForce = suspForce;
        if (suspForce != 0) {
            V3 off = wh.contactPoint - pos;
            float scale = (suspForce * dt) + c.wheelPush[i];
            V3 impulse = wh.contactNormal * scale;
            vel += impulse * k.invMass;
            angvel += iiw * cross(off, impulse);
        }
    }
    V3 upDir = c.rot.col(2);
    RL_WHEEL_LOOP
    for (int i = 0; i < 4; i++) {
        WheelW& wh = w.w[i];
        if (!is_zero(wh.impulse)) {
            V3 off = wh.contactPoint - pos;
            float upDot = dot(upDir, off);
            V3 rel = off - upDir * upDot;
            V3 impulse = wh.impulse * dt;
            vel += impulse * k.invMass;
            angvel += iiw * cross(rel, impulse);
        }
    }
    c.vel = vel; c.angvel = angvel;
}

// ---- Car::_UpdateBoost (Car.cpp:477-505) ------------------------------------------------------
RL_HD inline void update_boost(CarS& c, CarW& w, const SimCfg& cfg) {
    const float dt = kTickTime;
    if (c.timeSpentBoosting > 0) {
        if (!c.controls.boost && c.timeSpentBoosting >= C::BOOST_MIN_TIME) c.timeSpentBoosting = 0;
        else c.timeSpentBoosting += dt;
    } else if (c.controls.boost) {
        c.timeSpentBoosting = dt;
    }
    if (c.boost > 0 && c.timeSpentBoosting > 0) {
        c.boost = fmaxf_(c.boost - cfg.mut.boostUsedPerSecond * dt, 0.f);
        w.force += (c.isOnGround ? cfg.mut.boostAccelGround : cfg.mut.boostAccelAir) * UU2BT * c.rot.col(0) * C::CAR_MASS;
    }
    c.boost = fminf_(c.boost, C::BOOST_MAX);
}

// Car::SetState with a default CarState (Car.cpp:23-36, Car.h:17-101): wheel carry-over values and
// controls are NOT touched, exactly like the reference.
RL_HDI void car_set_default(CarS& c, float spawnBoost) {
    c.vel = V3(); c.angvel = V3();
    c.isOnGround = 1;
    for (int i = 0; i < 4; i++) c.wheelContact[i] = 0;
    c.hasJumped = c.hasDoubleJumped = c.hasFlipped = c.isFlipping = c.isJumping = 0;
    c.flipRelTorque = V3();
    c.jumpTime = c.flipTime = c.airTime = c.airTimeSinceJump = 0;
    c.boost = spawnBoost; c.timeSpentBoosting = 0;
    c.isSupersonic = 0; c.supersonicTime = 0; c.handbrakeVal = 0;
    c.isAutoFlipping = 0; c.autoFlipTimer = 0; c.autoFlipTorqueScale = 0;
    c.worldContactHas = 0; c.worldContactNormal = V3();
    c.carContactOtherId = 0; c.carContactCooldown = 0;
    c.isDemoed = 0; c.demoRespawnTimer = 0;
    c.hitValid = 0; c.hitRelPos = V3(); c.hitBallPos = V3(); c.hitExtraVel = V3();
    c.hitTickLo = c.hitTickHi = c.hitExtraTickLo = c.hitExtraTickHi = -1;
    c.lastControls = Controls{0, 0, 0, 0, 0, 0, 0, 0};
}
// Car::Respawn (Car.cpp:43-56)
RL_HD inline void car_respawn(CarS& c, int team, uint32_t rnd, float spawnBoost);

// ---- Car::_PreTickUpdate (Car.cpp:58-131) -----------------------------------------------------
// respawnRnd: a random word for Car::Respawn's spawn-slot pick; derived by the caller from the arena RNG state, the
// tick and the car index WITHOUT advancing the arena RNG (the roles of a tick run concurrently)
// query box of a car's mesh candidates: hitbox AABB united with the four wheel-ray segments, padded
RL_HDI void car_cands_box(const CarS& c, const CarConsts& k, V3& mnOut, V3& mxOut) {
    V3 center = c.pos + c.rot * k.hitboxOffset;
    V3 ext(dot(vabs(c.rot.r[0]), k.halfExt), dot(vabs(c.rot.r[1]), k.halfExt), dot(vabs(c.rot.r[2]), k.halfExt));
    V3 mn = center - ext, mx = center + ext;
    V3 wheelDir = c.rot * V3(0, 0, -1);
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        V3 hp = c.pos + c.rot * k.wheelConn[i];
        float rayLen = k.wheelRest[i] + k.suspTravel + k.wheelRadius[i] - C::SUSPENSION_SUBTRACTION;
        V3 tg = hp + wheelDir * rayLen;
        mn = vmin(mn, vmin(hp, tg)); mx = vmax(mx, vmax(hp, tg));
    }
    const V3 pad(0.02f, 0.02f, 0.02f);  // > the 0.01 box padding of the direct ray walk
    mnOut = mn - pad; mxOut = mx + pad;
}
// part A: up to the pose being final for this tick and (collect) the mesh candidates collected
RL_HD inline void car_pre_tick_a(CarS& c, const SimCfg& cfg, const MeshSet& ms, const CarConsts& k, int ci, CarW& w, uint32_t respawnRnd,
                                 bool collect = true) {
    w.force = V3(); w.torque = V3(); w.velCache = V3();
    c.controls.throttle = clampf(c.controls.throttle, -1.f, 1.f);
    c.controls.steer = clampf(c.controls.steer, -1.f, 1.f);
    c.controls.pitch = clampf(c.controls.pitch, -1.f, 1.f);
    c.controls.yaw = clampf(c.controls.yaw, -1.f, 1.f);
    c.controls.roll = clampf(c.controls.roll, -1.f, 1.f);
    if (c.isDemoed) {
        c.demoRespawnTimer = fmaxf_(c.demoRespawnTimer - kTickTime, 0.f);
        if (c.demoRespawnTimer == 0) car_respawn(c, car_team(ci, cfg.spawnOpponents), respawnRnd, cfg.mut.carSpawnBoost);
    }
    w.invInertiaWorld = world_inertia(c.rot, k.invInertiaLocal);
    if (collect) {  // one BVH query for the hitbox and the four wheel rays (pose is final for this tick: only a respawn moves it)
        V3 mn, mx;
        car_cands_box(c, k, mn, mx);
        collect_candidates(ms, mn, mx, w.cands);
    }
    RL_PT(0);
}
// part B: the vehicle update and the control -> force model (needs w.meshHit: wheel_mesh_rays / cands_pass_warp)
RL_HD inline void car_pre_tick_b(CarS& c, const TickX& x, const SimCfg& cfg, const MeshSet& ms, const CarConsts& k, int ci, CarW& w) {
    if (c.isDemoed) return;

    vehicle_first(c, x, cfg, ms, k, ci, w);
    RL_PT(1);
    bool jumpPressed = c.controls.jump && !c.lastControls.jump;
    int n = 0;
    for (int i = 0; i < 4; i++) { c.wheelContact[i] = w.w[i].inContact; n += w.w[i].inContact; }
    c.isOnGround = n >= 3;
    float forwardSpeedUU = dot(c.vel, c.rot.col(0)) * BT2UU;
    update_wheels(c, w, n, forwardSpeedUU);
    if (n < 3) update_air_torque(c, w, k, n == 0);
    else c.isFlipping = 0;
    update_jump(c, w, k, cfg, jumpPressed);
    update_auto_flip(c, k, jumpPressed);
    update_double_jump_or_flip(c, k, cfg, jumpPressed, forwardSpeedUU);
    if (c.controls.throttle != 0 && ((n > 0 && n < 4) || c.worldContactHas)) update_auto_roll(c, w, k, n);
    c.worldContactHas = 0;
    RL_PT(2);
    vehicle_second(c, w, k);
    update_boost(c, w, cfg);
    RL_PT(3);
}
RL_HD inline void car_pre_tick(CarS& c, const TickX& x, const SimCfg& cfg, const MeshSet& ms, const CarConsts& k, int ci, CarW& w, uint32_t respawnRnd) {
    car_pre_tick_a(c, cfg, ms, k, ci, w, respawnRnd);
    wheel_mesh_rays(c, k, ms, w);
    car_pre_tick_b(c, x, cfg, ms, k, ci, w);
}

RL_HD inline void car_respawn(CarS& c, int team, uint32_t rnd, float spawnBoost) {
    const float RX[4] = {-2304, -2688, 2304, 2688};
    const float RY = -4608;
    int idx = (int)(rnd % 4u);
    car_set_default(c, spawnBoost);
    V3 pos(RX[idx], RY * (team == 0 ? 1.f : -1.f), C::CAR_RESPAWN_Z);
    float yaw = (float)(3.14159265358979323846 / 2 + (team == 0 ? 0.0 : 3.14159265358979323846));
    c.pos = V3(pos.x * UU2BT, pos.y * UU2BT, pos.z * UU2BT);
    c.rot = angle_to_rotmat(yaw, 0.f, 0.f);
    c.boost = spawnBoost;
}

// ---- Car::_PostTickUpdate + _FinishPhysicsTick (Car.cpp:133-193) ------------------------------
RL_HD inline void car_post_tick(CarS& c, CarW& w) {
    if (c.isDemoed) return;
    float speedSq = len2(c.vel * BT2UU);
    if (c.isSupersonic && c.supersonicTime < C::SUPERSONIC_MAINTAIN_MAX_TIME)
        c.isSupersonic = speedSq >= C::SUPERSONIC_MAINTAIN_MIN_SPEED * C::SUPERSONIC_MAINTAIN_MIN_SPEED;
    else
        c.isSupersonic = speedSq >= C::SUPERSONIC_START_SPEED * C::SUPERSONIC_START_SPEED;
    if (c.isSupersonic) c.supersonicTime += kTickTime; else c.supersonicTime = 0;
    if (c.carContactCooldown > 0) c.carContactCooldown = fmaxf_(c.carContactCooldown - kTickTime, 0.f);
    c.lastControls = c.controls;
    // _FinishPhysicsTick
    if (!is_zero(w.velCache)) { c.vel += w.velCache; w.velCache = V3(); }
    const float maxSpeed = C::CAR_MAX_SPEED * UU2BT;
    if (len2(c.vel) > maxSpeed * maxSpeed) c.vel = normalized(c.vel) * maxSpeed;
    if (len2(c.angvel) > C::CAR_MAX_ANG_SPEED * C::CAR_MAX_ANG_SPEED) c.angvel = normalized(c.angvel) * C::CAR_MAX_ANG_SPEED;
}

}  // namespace rl
