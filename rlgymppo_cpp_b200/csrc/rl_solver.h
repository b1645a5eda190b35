// rl_solver.h — RocketSim's modified btSequentialImpulseConstraintSolver for one arena.
//
// Reference: B/BulletDynamics/ConstraintSolver/btSequentialImpulseConstraintSolver.cpp
//   setupContactConstraint :795-983, convertContact/Inner :1003-1162, convertContactSpecial :1164-1211,
//   solveSingleIteration :1601-1776, split impulse :1778-1829, writeBackBodies :1878-1904.
// Properties relied upon: manifolds are stateless (no warm start), one friction
// direction per contact chosen from the relative lateral velocity, restitution threshold 0.2,
// erp2 = 0.8, split impulse always, 10 iterations, RocketSim removed the penetration->velocity term,
// ball-world ("special") points are skipped in velocity iterations and replaced by ONE averaged
// contact per body, but still take part in the split-impulse position iterations.
#pragma once
#include "rl_collide.h"

namespace rl {

struct SolverBody {
    V3 pos;
    M3 rot;
    V3 linVel, angVel, extForceImp, extTorqueImp;
    V3 dLin, dAng, push, turn;
    M3 invInertiaWorld;
    float invMass;
    int32_t active;  // has an m_originalBody in an awake island
    // uninitialised on purpose (see NoInit): solver_body_from_car / tick_p2_solve write every member of an active body,
    // and an inactive body is never read past `active`
    RL_HDI SolverBody() : pos(NoInit()), rot(NoInit()), linVel(NoInit()), angVel(NoInit()), extForceImp(NoInit()), extTorqueImp(NoInit()),
                          dLin(NoInit()), dAng(NoInit()), push(NoInit()), turn(NoInit()), invInertiaWorld(NoInit()) {}
};

struct Row {
    V3 n1, rxn1, n2, rxn2, angA, angB;
    float jacDiagInv, rhs, rhsPen, lower, upper, applied, appliedPush, friction;
    int32_t a, b;        // solver body indices; -1 = fixed body
    int32_t frictionIndex, special;
    // uninitialised on purpose (see NoInit): setup_contact_row / setup_friction_row write every member
    RL_HDI Row() : n1(NoInit()), rxn1(NoInit()), n2(NoInit()), rxn2(NoInit()), angA(NoInit()), angB(NoInit()) {}
};

constexpr int kMaxRows = kMaxContacts + kMaxCars + 1;

RL_HDI void sb_apply(SolverBody* sb, int i, V3 lin, V3 ang, float mag) {
    if (i < 0) return;
    sb[i].dLin += lin * mag;
    sb[i].dAng += ang * mag;
}
RL_HDI void sb_apply_push(SolverBody* sb, int i, V3 lin, V3 ang, float mag) {
    if (i < 0) return;
    sb[i].push += lin * mag;
    sb[i].turn += ang * mag;
}

// btPlaneSpace1
RL_HDI void plane_space(V3 n, V3& p, V3& q) {
    if (fabsf(n.z) > 0.7071067811865475244008443621048490f) {
        float a = n.y * n.y + n.z * n.z;
        float k = 1.f / sqrtf(a);
        p = V3(0, -n.z * k, n.y * k);
        q = V3(a * k, -n.x * p.z, n.x * p.y);
    } else {
        float a = n.x * n.x + n.y * n.y;
        float k = 1.f / sqrtf(a);
        p = V3(-n.y * k, n.x * k, 0);
        q = V3(-n.z * p.y, n.z * p.x, a * k);
    }
}

RL_HD RL_NOINLINE inline void setup_contact_row(Row& r, const SolverBody* sb, int ia, int ib, V3 n, V3 rel1, V3 rel2, float dist, float restitutionCoef, float frictionCoef, int special) {
    const float invDt = 1.f / kTickTime;
    bool hasA = ia >= 0, hasB = ib >= 0;
    r.a = ia; r.b = ib;
    V3 torqueAxis0 = cross(rel1, n);
    V3 torqueAxis1 = cross(rel2, n);
    r.angA = hasA ? sb[ia].invInertiaWorld * torqueAxis0 : V3();
    r.angB = hasB ? sb[ib].invInertiaWorld * (-torqueAxis1) : V3();
    float denom0 = 0.f, denom1 = 0.f;
    if (hasA) denom0 = sb[ia].invMass + dot(n, cross(r.angA, rel1));
    if (hasB) denom1 = sb[ib].invMass + dot(n, cross(-r.angB, rel2));
    r.jacDiagInv = 1.f / (denom0 + denom1 + 0.f);
    r.n1 = hasA ? n : V3(); r.rxn1 = hasA ? torqueAxis0 : V3();
    r.n2 = hasB ? -n : V3(); r.rxn2 = hasB ? -torqueAxis1 : V3();
    float penetration = dist + 0.f;
    V3 vel1 = hasA ? vel_at(sb[ia].linVel, sb[ia].angVel, rel1) : V3();
    V3 vel2 = hasB ? vel_at(sb[ib].linVel, sb[ib].angVel, rel2) : V3();
    float rel_vel = dot(n, vel1 - vel2);
    r.friction = frictionCoef;
    float restitution = fabsf(rel_vel) < C::RESTITUTION_VEL_THRESH ? 0.f : restitutionCoef * -rel_vel;
    if (restitution <= 0.f) restitution = 0.f;
    r.applied = 0.f; r.appliedPush = 0.f;
    V3 efA = hasA ? sb[ia].extForceImp : V3(), etA = hasA ? sb[ia].extTorqueImp : V3();
    V3 efB = hasB ? sb[ib].extForceImp : V3(), etB = hasB ? sb[ib].extTorqueImp : V3();
    V3 lvA = hasA ? sb[ia].linVel : V3(), avA = hasA ? sb[ia].angVel : V3();
    V3 lvB = hasB ? sb[ib].linVel : V3(), avB = hasB ? sb[ib].angVel : V3();
    float vel1Dotn = dot(r.n1, lvA + efA) + dot(r.rxn1, avA + etA);
    float vel2Dotn = dot(r.n2, lvB + efB) + dot(r.rxn2, avB + etB);
    float rv = vel1Dotn + vel2Dotn;
    float positionalError = 0.f;
    float velocityError = restitution - rv;
    if (penetration > 0) positionalError = 0;
    else positionalError = -penetration * C::ERP2 * invDt;
    r.rhs = velocityError * r.jacDiagInv;
    r.rhsPen = positionalError * r.jacDiagInv;
    r.lower = 0; r.upper = 1e10f;
    r.special = special;
    r.frictionIndex = -1;
}

RL_HD RL_NOINLINE inline void setup_friction_row(Row& r, const SolverBody* sb, int ia, int ib, V3 axis, V3 rel1, V3 rel2, float friction, int contactIndex) {
    bool hasA = ia >= 0, hasB = ib >= 0;
    r.a = ia; r.b = ib;
    r.friction = friction; r.applied = 0.f; r.appliedPush = 0.f;
    if (hasA) { r.n1 = axis; r.rxn1 = cross(rel1, r.n1); r.angA = sb[ia].invInertiaWorld * r.rxn1; }
    else { r.n1 = V3(); r.rxn1 = V3(); r.angA = V3(); }
    if (hasB) { r.n2 = -axis; r.rxn2 = cross(rel2, r.n2); r.angB = sb[ib].invInertiaWorld * r.rxn2; }
    else { r.n2 = V3(); r.rxn2 = V3(); r.angB = V3(); }
    float denom0 = 0.f, denom1 = 0.f;
    if (hasA) denom0 = sb[ia].invMass + dot(axis, cross(r.angA, rel1));
    if (hasB) denom1 = sb[ib].invMass + dot(axis, cross(-r.angB, rel2));
    r.jacDiagInv = 1.f / (denom0 + denom1);
    float vel1Dotn = dot(r.n1, hasA ? sb[ia].linVel + sb[ia].extForceImp : V3()) + dot(r.rxn1, hasA ? sb[ia].angVel : V3());
    float vel2Dotn = dot(r.n2, hasB ? sb[ib].linVel + sb[ib].extForceImp : V3()) + dot(r.rxn2, hasB ? sb[ib].angVel : V3());
    float rel_vel = vel1Dotn + vel2Dotn;
    r.rhs = (0.f - rel_vel) * r.jacDiagInv;
    r.rhsPen = 0.f;
    r.lower = -friction; r.upper = friction;
    r.frictionIndex = contactIndex; r.special = 0;
}

RL_HDI V3 vel_no_delta(const SolverBody* sb, int i, V3 rel) {
    if (i < 0) return V3();
    return sb[i].linVel + sb[i].extForceImp + cross(sb[i].angVel + sb[i].extTorqueImp, rel);
}

RL_HDI float resolve_row(SolverBody* sb, Row& c, bool lowerOnly) {
    float deltaImpulse = c.rhs - c.applied * 0.f;
    float dv1 = c.a >= 0 ? dot(c.n1, sb[c.a].dLin) + dot(c.rxn1, sb[c.a].dAng) : 0.f;
    float dv2 = c.b >= 0 ? dot(c.n2, sb[c.b].dLin) + dot(c.rxn2, sb[c.b].dAng) : 0.f;
    deltaImpulse -= dv1 * c.jacDiagInv;
    deltaImpulse -= dv2 * c.jacDiagInv;
    float sum = c.applied + deltaImpulse;
    if (sum < c.lower) { deltaImpulse = c.lower - c.applied; c.applied = c.lower; }
    else if (!lowerOnly && sum > c.upper) { deltaImpulse = c.upper - c.applied; c.applied = c.upper; }
    else c.applied = sum;
    if (c.a >= 0) sb_apply(sb, c.a, c.n1 * sb[c.a].invMass, c.angA, deltaImpulse);
    if (c.b >= 0) sb_apply(sb, c.b, c.n2 * sb[c.b].invMass, c.angB, deltaImpulse);
    return deltaImpulse;
}

RL_HDI float resolve_split(SolverBody* sb, Row& c) {
    float deltaImpulse = 0.f;
    if (c.rhsPen != 0.f) {
        deltaImpulse = c.rhsPen - c.appliedPush * 0.f;
        float dv1 = c.a >= 0 ? dot(c.n1, sb[c.a].push) + dot(c.rxn1, sb[c.a].turn) : 0.f;
        float dv2 = c.b >= 0 ? dot(c.n2, sb[c.b].push) + dot(c.rxn2, sb[c.b].turn) : 0.f;
        deltaImpulse -= dv1 * c.jacDiagInv;
        deltaImpulse -= dv2 * c.jacDiagInv;
        float sum = c.appliedPush + deltaImpulse;
        if (sum < c.lower) { deltaImpulse = c.lower - c.appliedPush; c.appliedPush = c.lower; }
        else c.appliedPush = sum;
        if (c.a >= 0) sb_apply_push(sb, c.a, c.n1 * sb[c.a].invMass, c.angA, deltaImpulse);
        if (c.b >= 0) sb_apply_push(sb, c.b, c.n2 * sb[c.b].invMass, c.angB, deltaImpulse);
    }
    return deltaImpulse;
}

// btTransformUtil::integrateTransform (B/LinearMath/btTransformUtil.h:37-88)
RL_HD inline void integrate_transform(V3& pos, M3& rot, V3 linvel, V3 angvel, float dt) {
    pos = pos + linvel * dt;
    float fAngle2 = len2(angvel);
    float fAngle = 0;
    if (fAngle2 > kEps) fAngle = sqrtf(fAngle2);
    const float ANGULAR_MOTION_THRESHOLD = 0.5f * kHalfPi;
    if (fAngle * dt > ANGULAR_MOTION_THRESHOLD) fAngle = ANGULAR_MOTION_THRESHOLD / dt;
    V3 axis;
    if (fAngle < 0.001f) axis = angvel * (0.5f * dt - (dt * dt * dt) * 0.020833333333f * fAngle * fAngle);
    else axis = angvel * (rl_sin(0.5f * fAngle * dt) / fAngle);
    Quat dorn(axis.x, axis.y, axis.z, rl_cos(fAngle * dt * 0.5f));
    Quat orn0 = mat_to_quat(rot);
    Quat pred = dorn * orn0;
    float l2 = pred.x * pred.x + pred.y * pred.y + pred.z * pred.z + pred.w * pred.w;
    if (l2 > kEps * kEps) { float s = 1.f / sqrtf(l2); pred = Quat(pred.x * s, pred.y * s, pred.z * s, pred.w * s); }  // safeNormalize
    else pred = Quat(1, 0, 0, 0);
    float pl2 = pred.x * pred.x + pred.y * pred.y + pred.z * pred.z + pred.w * pred.w;
    if (pl2 > kEps) rot = quat_to_mat(pred);
}

// solveGroup for one simulation island.  `sb` holds the island's bodies; a contact's body index i (0 ball, 1+c car c,
// -1 static) maps to sb[i - base].  Islands that share no body do not interact in a Gauss-Seidel sweep, so solving them
// one by one (each with its rows in the reference's manifold order) gives exactly what the reference's single
// solveGroup over all manifolds gives; this is what lets every role of a tick solve its own island concurrently
// (rl_tick.h).  The split-impulse early exit on a zero residual is per island for the same reason: an island whose
// rows all returned a zero delta is at a fixed point of the sweep.
// rows of one island in the reference's order: a contact row + its friction row per contact, then the averaged special rows
RL_HD RL_NOINLINE inline void island_setup(const SolverBody* sb, int numBodies, int base, const Contact* contacts, int numContacts, Row* rows, Row* fric,
                                          int& nRowsOut, int& nFricOut) {
    int nRows = 0, nFric = 0;
    // special-contact accumulators per body (btCollisionObject::m_specialResolveInfo)
    int spN[1 + kMaxCars]; float spFriction[1 + kMaxCars], spRestitution[1 + kMaxCars], spDist[1 + kMaxCars]; V3 spNormal[1 + kMaxCars];
    for (int i = 0; i < numBodies; i++) { spN[i] = 0; spDist[i] = 0; spNormal[i] = V3(); spFriction[i] = spRestitution[i] = 0; }

    for (int ci = 0; ci < numContacts; ci++) {
        const Contact cp = contacts[ci];  // one 64-byte fetch instead of field-by-field round trips to the scratch segment
        const int ia = cp.a < 0 ? -1 : cp.a - base, ib = cp.b < 0 ? -1 : cp.b - base;
        // a body without contact response (demoed car) or in a sleeping island (frozen ball) takes its manifolds out of
        // the solver entirely (btCollisionDispatcher::needsResponse / island filtering); it is NOT a static obstacle
        if ((ia >= 0 && !sb[ia].active) || (ib >= 0 && !sb[ib].active)) continue;
        if (nRows >= kMaxRows - (1 + kMaxCars)) break;
        V3 rel1 = cp.posA - (ia >= 0 ? sb[ia].pos : V3());
        V3 rel2 = cp.posB - (ib >= 0 ? sb[ib].pos : V3());
        int idx = nRows;
        setup_contact_row(rows[nRows], sb, ia, ib, cp.normal, rel1, rel2, cp.dist, cp.restitution, cp.friction, cp.special);
        rows[nRows].frictionIndex = nFric;
        nRows++;
        if (cp.special) {
            for (int s = 0; s < 2; s++) {
                int bi = s ? ib : ia;
                if (bi >= 0) {
                    spN[bi]++; spFriction[bi] = cp.friction; spRestitution[bi] = cp.restitution;
                    spNormal[bi] += cp.normal; spDist[bi] += len(s ? rel2 : rel1);
                }
            }
        }
        // convertContactInner: one friction direction from the lateral relative velocity
        V3 vel = vel_no_delta(sb, ia, rel1) - vel_no_delta(sb, ib, rel2);
        float rel_vel = dot(cp.normal, vel);
        V3 lat = vel - cp.normal * rel_vel;
        float lat2 = len2(lat);
        V3 dir, dir2;
        if (lat2 > kEps) dir = lat * (1.f / sqrtf(lat2));
        else plane_space(cp.normal, dir, dir2);
        setup_friction_row(fric[nFric], sb, ia, ib, dir, rel1, rel2, cp.friction, idx);
        nFric++;
    }
    // convertContactSpecial: one averaged contact per body with special collisions
    for (int bi = 0; bi < numBodies; bi++) {
        if (spN[bi] == 0 || !sb[bi].active) continue;
        float distance = spDist[bi] / (float)spN[bi];
        V3 normal = spNormal[bi] / (float)spN[bi];
        V3 rel1 = normal * -distance;
        int idx = nRows;
        setup_contact_row(rows[nRows], sb, bi, -1, normal, rel1, V3(), distance, spRestitution[bi], spFriction[bi], 0);
        rows[nRows].frictionIndex = nFric;
        nRows++;
        V3 vel = vel_no_delta(sb, bi, rel1);
        float rel_vel = dot(normal, vel);
        V3 lat = vel - normal * rel_vel;
        float lat2 = len2(lat);
        V3 dir, dir2;
        if (lat2 > kEps) dir = lat * (1.f / sqrtf(lat2));
        else plane_space(normal, dir, dir2);
        setup_friction_row(fric[nFric], sb, bi, -1, dir, rel1, V3(), spFriction[bi], idx);
        nFric++;
    }
    nRowsOut = nRows; nFricOut = nFric;
}

// writeBackBodies (+ the split-impulse transform correction)
RL_HD RL_NOINLINE inline void island_finish(SolverBody* sb, int numBodies) {
    for (int i = 0; i < numBodies; i++) {
        if (!sb[i].active) continue;
        sb[i].linVel += sb[i].dLin;
        sb[i].angVel += sb[i].dAng;
        if (!is_zero(sb[i].push) || !is_zero(sb[i].turn)) integrate_transform(sb[i].pos, sb[i].rot, sb[i].push, sb[i].turn * 0.1f, kTickTime);
        sb[i].linVel = sb[i].linVel + sb[i].extForceImp;
        sb[i].angVel = sb[i].angVel + sb[i].extTorqueImp;
    }
}

RL_HD RL_NOINLINE inline void solve_island(SolverBody* sb, int numBodies, int base, const Contact* contacts, int numContacts) {
    Row rows[kMaxRows];
    Row fric[kMaxRows];
    int nRows, nFric;
    island_setup(sb, numBodies, base, contacts, numContacts, rows, fric, nRows, nFric);

    const int numIterations = 10;
    // split impulse (position) iterations — includes the special rows
    for (int it = 0; it < numIterations; it++) {
        float residual = 0.f;
        for (int j = 0; j < nRows; j++) {
            float d = resolve_split(sb, rows[j]) * (1.f / rows[j].jacDiagInv);
            residual = fmaxf_(residual, d * d);
        }
        if (residual <= 0.f || it >= numIterations - 1) break;
    }
    for (int it = 0; it < numIterations; it++) {
        for (int j = 0; j < nRows; j++) {
            if (rows[j].special) continue;
            resolve_row(sb, rows[j], true);
        }
        for (int j = 0; j < nFric; j++) {
            float total = rows[fric[j].frictionIndex].applied;
            if (total > 0.f) {
                fric[j].lower = -(fric[j].friction * total);
                fric[j].upper = fric[j].friction * total;
                resolve_row(sb, fric[j], false);
            }
        }
    }
    island_finish(sb, numBodies);
}

// The common island: ONE dynamic body against the static world (the ball on the floor / a wall, a car's hitbox on the
// ground) — every contact has a == body, b == -1.  Same rows and the same sweep as solve_island, but the body's
// velocity / push deltas live in registers for the 10 + 10 iterations instead of round-tripping through the
// SolverBody array between rows (the B side of every row is the fixed body: its terms are exact zeros).
RL_HD RL_NOINLINE inline void solve_island_one(SolverBody& b, int body, const Contact* contacts, int numContacts) {
    Row rows[kMaxRows];
    Row fric[kMaxRows];
    int nRows, nFric;
    RL_PT(-1);
    island_setup(&b, 1, body, contacts, numContacts, rows, fric, nRows, nFric);
    RL_PT(23);
    if (b.active) {
        const float invMass = b.invMass;
        V3 dLin = b.dLin, dAng = b.dAng, push = b.push, turn = b.turn;
        const int numIterations = 10;
        for (int it = 0; it < numIterations; it++) {
            float residual = 0.f;
            for (int j = 0; j < nRows; j++) {
                Row& c = rows[j];
                float deltaImpulse = 0.f;
                if (c.rhsPen != 0.f) {  // resolve_split
                    deltaImpulse = c.rhsPen - c.appliedPush * 0.f;
                    float dv1 = dot(c.n1, push) + dot(c.rxn1, turn);
                    deltaImpulse -= dv1 * c.jacDiagInv;
                    float sum = c.appliedPush + deltaImpulse;
                    if (sum < c.lower) { deltaImpulse = c.lower - c.appliedPush; c.appliedPush = c.lower; }
                    else c.appliedPush = sum;
                    push += (c.n1 * invMass) * deltaImpulse;
                    turn += c.angA * deltaImpulse;
                }
                float d = deltaImpulse * (1.f / c.jacDiagInv);
                residual = fmaxf_(residual, d * d);
            }
            if (residual <= 0.f || it >= numIterations - 1) break;
        }
        RL_PT(24);
        auto resolve = [&](Row& c, bool lowerOnly) {  // resolve_row
            float deltaImpulse = c.rhs - c.applied * 0.f;
            float dv1 = dot(c.n1, dLin) + dot(c.rxn1, dAng);
            deltaImpulse -= dv1 * c.jacDiagInv;
            float sum = c.applied + deltaImpulse;
            if (sum < c.lower) { deltaImpulse = c.lower - c.applied; c.applied = c.lower; }
            else if (!lowerOnly && sum > c.upper) { deltaImpulse = c.upper - c.applied; c.applied = c.upper; }
            else c.applied = sum;
            dLin += (c.n1 * invMass) * deltaImpulse;
            dAng += c.angA * deltaImpulse;
        };
        for (int it = 0; it < numIterations; it++) {
            for (int j = 0; j < nRows; j++) {
                if (rows[j].special) continue;
                resolve(rows[j], true);
            }
            for (int j = 0; j < nFric; j++) {
                float total = rows[fric[j].frictionIndex].applied;
                if (total > 0.f) {
                    fric[j].lower = -(fric[j].friction * total);
                    fric[j].upper = fric[j].friction * total;
                    resolve(fric[j], false);
                }
            }
        }
        b.dLin = dLin; b.dAng = dAng; b.push = push; b.turn = turn;
    }
    RL_PT(25);
    island_finish(&b, 1);
    RL_PT(26);
}

}  // namespace rl
