// rl_collide.h — narrowphase for the shape pairs RocketSim's soccar actually reaches
// (SURVEY.md §8a "Narrowphase pairs actually reached") plus the contact-added callback
// (R/Sim/Arena/Arena.cpp:218-427).  Contacts are stateless per tick: the reference's
// btRSBroadphase drops and rebuilds every pair (and therefore every manifold) each tick.
//
// Body indices: 0 = ball, 1+c = car c, -1 = static world.  A manifold's body A is the
// dynamic one for world pairs, the car for car-ball, the HIGHER car for car-car (see car_car); normals
// are "normalWorldOnB": they point from B towards A.
#pragma once
#include "rl_car.h"
#include "rl_gjk.h"

namespace rl {

struct alignas(16) Contact {  // 64 bytes: moves between the scratch segments (global memory) and the roles as 4 x 16 B
    int32_t a, b;
    V3 posA, posB, normal;
    float dist, friction, restitution;
    int32_t special;
    int32_t pad_;
    // uninitialised on purpose (see NoInit): manifold_add writes every member before a contact is stored
    RL_HDI Contact() : posA(NoInit()), posB(NoInit()), normal(NoInit()) {}
};

constexpr int kMaxContacts = 40;

struct Manifold {
    int32_t a, b, n;
    float breaking;
    Contact pt[4];
};

struct ContactSet {
    int32_t n;
    int32_t overflow;
    Contact c[kMaxContacts];
};

// Where a role appends the contacts it finds: a segment of the arena's contact scratch (global memory on the device).
// Segment layout per arena: [ball: kSegBall][per car: 1 car-ball slot + kSegCarWorld world slots + kSegCarPlane staging slots][car-car: kSegPair]
// A car's world contacts are produced in two pieces that may be written by different roles at the same time (rl_tick.h): the
// hitbox-mesh contacts fill the world slots from the front, the (at most 4) hitbox-plane contacts go to the staging slots; readers
// see "mesh piece, then plane piece", capped at kSegCarWorld in total — the order and the cap of one shared sink.
constexpr int kSegBall = 12, kSegCarWorld = 13, kSegCarPlane = 4, kSegCar = 1 + kSegCarWorld + kSegCarPlane, kSegPair = 16;
RL_HDI int contact_scratch_slots(int ncars) { return kSegBall + ncars * kSegCar + kSegPair; }
struct ContactSink {
    Contact* base;
    int32_t n, cap, overflow;
};
RL_HDI ContactSink make_sink(Contact* base, int cap) { ContactSink s; s.base = base; s.n = 0; s.cap = cap; s.overflow = 0; return s; }

// relative contact breaking thresholds (btCollisionDispatcher::getNewManifold,
// btCollisionShape::getContactBreakingThreshold): 0.02 * angularMotionDisc of the shape
struct Thresholds { float ball, car; };
RL_HDI Thresholds contact_thresholds(const CarConsts& k, float ballRadiusUU = C::BALL_RADIUS) {
    Thresholds t;
    // btCollisionShape::getBoundingSphere, sphere case (ROCKETSIM CHANGE: radius + 0.08), centre 0
    t.ball = (float)((double)(ballRadiusUU * UU2BT) + 0.08) * C::CONTACT_BREAKING;
    // compound: local AABB = hitbox offset +- half extents
    V3 mn = k.hitboxOffset - k.halfExt, mx = k.hitboxOffset + k.halfExt;
    float radius = len(mx - mn) * 0.5f;
    V3 center = (mn + mx) * 0.5f;
    t.car = (radius + len(center)) * C::CONTACT_BREAKING;
    return t;
}

// btPersistentManifold::sortCachedPoints with gContactCalcArea3Points (btPersistentManifold.cpp:110-190)
RL_HD inline int manifold_sort_cached(const Manifold& m, const Contact& pt) {
    int maxPenIdx = -1;
    float maxPen = pt.dist;
    for (int i = 0; i < 4; i++)
        if (m.pt[i].dist < maxPen) { maxPenIdx = i; maxPen = m.pt[i].dist; }
    float res[4] = {0, 0, 0, 0};
    const V3 &p0 = m.pt[0].posA, &p1 = m.pt[1].posA, &p2 = m.pt[2].posA, &p3 = m.pt[3].posA;
    if (maxPenIdx != 0) res[0] = len2(cross(pt.posA - p1, p3 - p2));
    if (maxPenIdx != 1) res[1] = len2(cross(pt.posA - p0, p3 - p2));
    if (maxPenIdx != 2) res[2] = len2(cross(pt.posA - p0, p3 - p1));
    if (maxPenIdx != 3) res[3] = len2(cross(pt.posA - p0, p2 - p1));
    // btVector4::closestAxis4 == index of max |.|
    int best = 0; float bv = fabsf(res[0]);
    for (int i = 1; i < 4; i++) if (fabsf(res[i]) > bv) { bv = fabsf(res[i]); best = i; }
    return best;
}

struct CollideCtx {
    ArenaS* a;
    const SimCfg* cfg;
    TickX tx;   // tx.car[c].noResponse: DISABLE_SIMULATION | CF_NO_CONTACT_RESPONSE for this tick (demoed when it started, Car.cpp:69-87)
    const CarConsts* k;
    int64_t tick;
    V3 ballPos, ballVel;      // the ball as the narrowphase sees it: start-of-tick position, DAMPED velocity
    int32_t firstTickOfStep;  // bump counters only stick when the callback fires during Gym::Step's first tick (see rl_tick.h)
    const EpaCtx* epa;               // penetration-depth workspace (device: the warp's; host: nullptr = local, rl_epa.h)
    // car-world callback target: nullptr = the car's own CarState::worldContact; the role kernel's ball warp, which evaluates the
    // hitbox-mesh pairs while the car's own role is still inside Car::_PreTickUpdate (that reads and clears worldContact), records
    // the callback's result here instead and the car's role applies it afterwards (engine.cu, rl_tick.h car_world_merge)
    int32_t* wcHas = nullptr; V3* wcNormal = nullptr;
};

// ---- Arena::_BulletContactAddedCallback ------------------------------------------------------
RL_HD inline void on_car_ball(CollideCtx& x, int ci, Contact& cp) {
    ArenaS& a = *x.a;
    CarS& car = a.cars[ci];
    cp.friction = C::CARBALL_FRICTION; cp.restitution = C::CARBALL_RESTITUTION;
    V3 ballPosUU = to_uu(x.ballPos), ballVelUU = to_uu(x.ballVel);
    V3 carPosUU = to_uu(car.pos), carVelUU = to_uu(car.vel);
    car.hitValid = 1;
    car.hitRelPos = (cp.posB - x.ballPos) * BT2UU;  // m_localPointB of the ball (identity basis)
    set_i64(car.hitTickLo, car.hitTickHi, x.tick);
    car.hitBallPos = ballPosUU;
    car.hitExtraVel = V3();
    int64_t extraTick = get_i64(car.hitExtraTickLo, car.hitExtraTickHi);
    // uint64 compare: ~0ULL (== -1 here) means "never"
    bool never = extraTick < 0;
    if (never || (x.tick > extraTick + 1) || (extraTick > x.tick)) set_i64(car.hitExtraTickLo, car.hitExtraTickHi, x.tick);
    else return;
    V3 carForward = car.rot.col(0);
    V3 relPos = ballPosUU - carPosUU;
    V3 relVel = ballVelUU - carVelUU;
    float relSpeed = fminf_(len(relVel), C::BALL_CAR_EXTRA_IMPULSE_MAXDELTAVEL_UU);
    if (relSpeed > 0) {
        V3 hitDir = safe_normalized(relPos * V3(1, 1, C::BALL_CAR_EXTRA_IMPULSE_Z_SCALE));
        V3 fwdAdj = carForward * dot(hitDir, carForward) * (1 - C::BALL_CAR_EXTRA_IMPULSE_FORWARD_SCALE);
        hitDir = safe_normalized(hitDir - fwdAdj);
        const float fx[4] = {0, 500.f, 2300.f, 4600.f}, fy[4] = {0.65f, 0.65f, 0.55f, 0.30f};
        V3 addedVel = (hitDir * relSpeed) * curve(fx, fy, relSpeed) * x.cfg->mut.ballHitExtraForceScale;
        car.hitExtraVel = addedVel;
        x.tx.car[ci].ballVelCache += addedVel * UU2BT;
    }
}

RL_HD inline void on_car_car(CollideCtx& x, int c1, int c2, Contact& cp) {
    ArenaS& a = *x.a;
    cp.friction = C::CARCAR_FRICTION; cp.restitution = C::CARCAR_RESTITUTION;
    for (int i = 0; i < 2; i++) {
        bool swapped = i == 1;
        if (swapped) { int t = c1; c1 = c2; c2 = t; }
        CarS& s = a.cars[c1];
        CarS& o = a.cars[c2];
        if (s.isDemoed || o.isDemoed) return;
        if (s.carContactOtherId == c2 + 1 && s.carContactCooldown > 0) continue;
        V3 sPos = to_uu(s.pos), sVel = to_uu(s.vel), oPos = to_uu(o.pos), oVel = to_uu(o.vel);
        V3 deltaPos = oPos - sPos;
        if (ref_dot(sVel, deltaPos) > 0) {
            V3 velDir = ref_normalized(sVel);
            V3 dirToOther = ref_normalized(deltaPos);
            float speedTowards = ref_dot(sVel, dirToOther);
            float otherAway = ref_dot(oVel, velDir);
            if (speedTowards > otherAway) {
                // m_localPointA / m_localPointB in the respective car's frame (A = the higher car, see car_car)
                V3 wp = swapped ? cp.posB : cp.posA;
                const CarS& own = a.cars[c1];
                V3 local = tmul(wp - own.pos, own.rot);
                bool bumper = (local.x * BT2UU) > C::BUMP_MIN_FORWARD_DIST;
                if (bumper) {
                    const Mut& mu = x.cfg->mut;
                    bool isDemo = mu.demoMode == RLG_DEMO_ON_CONTACT ? true : (mu.demoMode == RLG_DEMO_DISABLED ? false : s.isSupersonic != 0);  // Arena.cpp:375-385
                    if (isDemo && !mu.enableTeamDemos) isDemo = car_team(c1, x.cfg->spawnOpponents) != car_team(c2, x.cfg->spawnOpponents);
                    if (isDemo) {
                        o.isDemoed = 1; o.demoRespawnTimer = mu.respawnDelay;
                    } else {
                        bool groundHit = o.isOnGround != 0;
                        const float gx[3] = {0.f, 1400.f, 2200.f}, gy[3] = {5.f / 6.f, 1100.f, 1530.f};
                        const float ax[3] = {0.f, 1400.f, 2200.f}, ay[3] = {5.f / 6.f, 1390.f, 1945.f};
                        const float ux[3] = {0.f, 1400.f, 2200.f}, uy[3] = {2.f / 6.f, 278.f, 417.f};
                        float baseScale = groundHit ? curve(gx, gy, speedTowards) : curve(ax, ay, speedTowards);
                        V3 hitUp = o.isOnGround ? o.rot.col(2) : V3(0, 0, 1);
                        V3 bump = velDir * baseScale + hitUp * curve(ux, uy, speedTowards) * mu.bumpForceScale;
                        x.tx.car[c2].velCache += bump * UU2BT;
                    }
                    s.carContactOtherId = c2 + 1;
                    s.carContactCooldown = mu.bumpCooldownTime;
                    // Gym's _BumpCallback (G/Gym.cpp:27-36): opponents only
                    if (x.firstTickOfStep && car_team(c1, x.cfg->spawnOpponents) != car_team(c2, x.cfg->spawnOpponents)) {
                        s.matchBumps++;
                        if (isDemo) s.matchDemos++;
                    }
                }
            }
        }
    }
}

RL_HDI void on_car_world(CollideCtx& x, int ci, Contact& cp) {
    if (x.wcHas) { *x.wcHas = 1; *x.wcNormal = cp.normal; }
    else {
        CarS& car = x.a->cars[ci];
        car.worldContactHas = 1;
        car.worldContactNormal = cp.normal;
    }
    cp.friction = x.cfg->mut.carWorldFriction; cp.restitution = x.cfg->mut.carWorldRestitution;
}

// btManifoldResult::addContactPoint (btManifoldResult.cpp:110-215) incl. the callback dispatch
RL_HD inline void manifold_add(CollideCtx& x, Manifold& m, V3 normalOnB, V3 pointOnB, float depth, const MeshSet* ms, int tri) {
    if (depth > m.breaking) return;
    Contact cp;
    cp.a = m.a; cp.b = m.b;
    cp.posA = pointOnB + normalOnB * depth;
    cp.posB = pointOnB;
    cp.normal = normalOnB;
    cp.dist = depth;
    cp.special = 0;
    // combined material (btManifoldResult.cpp:63-82): min friction / max restitution against statics, product otherwise
    if (m.a == 0 && m.b == -1) { cp.friction = fminf_(x.cfg->mut.ballWorldFriction, C::WORLD_FRICTION); cp.restitution = fmaxf_(x.cfg->mut.ballWorldRestitution, C::WORLD_RESTITUTION); }
    else { cp.friction = 0.3f; cp.restitution = 0.1f; }
    int idx = m.n;
    if (idx == 4) idx = manifold_sort_cached(m, cp);
    else m.n++;
    if (idx < 0) idx = 0;
    m.pt[idx] = cp;
    Contact& p = m.pt[idx];
    // gContactAddedCallback; demoed cars have no contact response -> returns before anything
    bool aCar = m.a >= 1, bCar = m.b >= 1;
    if (aCar && x.tx.car[m.a - 1].noResponse) return;
    if (bCar && x.tx.car[m.b - 1].noResponse) return;
    if (aCar && m.b == 0) on_car_ball(x, m.a - 1, p);
    else if (aCar && bCar) on_car_car(x, m.a - 1, m.b - 1, p);
    else if (aCar && m.b == -1) on_car_world(x, m.a - 1, p);
    else if (m.a == 0 && m.b == -1) p.special = 1;
    if (ms && tri >= 0) adjust_internal_edge(p, *ms, tri);
}

RL_HD RL_NOINLINE inline void manifold_flush(ContactSink& cs, const Manifold& m) {
    for (int i = 0; i < m.n; i++) {
        if (cs.n < cs.cap) cs.base[cs.n++] = m.pt[i];
        else cs.overflow++;
    }
}

// ---- shape pairs -----------------------------------------------------------------------------
// btConvexPlaneCollisionAlgorithm::processCollision (btConvexPlaneCollisionAlgorithm.cpp:92-125)
RL_HD inline void sphere_plane(CollideCtx& x, ContactSink& cs, V3 center, float radius, int planeIdx, float breaking) {
    PlaneDef p = world_plane(planeIdx);
    V3 cIn = center - p.origin;
    // localGetSupportingVertex(-n) = -n * radius (btSphereShape)
    V3 vtx = cIn + (-p.n) * radius;
    float distance = dot(p.n, vtx) - 0.f;
    if (distance < breaking) {
        Manifold m; m.a = 0; m.b = -1; m.n = 0; m.breaking = breaking;
        V3 proj = vtx - p.n * distance;
        manifold_add(x, m, p.n, proj + p.origin, distance, nullptr, -1);
        manifold_flush(cs, m);
    }
}

RL_HD inline void box_plane(CollideCtx& x, ContactSink& cs, int ci, int planeIdx, float breaking) {
    const CarS& c = x.a->cars[ci];
    const CarConsts& k = *x.k;
    PlaneDef p = world_plane(planeIdx);
    V3 boxCenter = c.pos + c.rot * k.hitboxOffset;
    V3 dirLocal = tmul(-p.n, c.rot);  // planeInConvex basis * -n
    V3 vl(dirLocal.x >= 0 ? k.halfExt.x : -k.halfExt.x, dirLocal.y >= 0 ? k.halfExt.y : -k.halfExt.y, dirLocal.z >= 0 ? k.halfExt.z : -k.halfExt.z);
    V3 vtx = (boxCenter + c.rot * vl) - p.origin;
    float distance = dot(p.n, vtx);
    if (distance < breaking) {
        Manifold m; m.a = 1 + ci; m.b = -1; m.n = 0; m.breaking = breaking;
        V3 proj = vtx - p.n * distance;
        manifold_add(x, m, p.n, proj + p.origin, distance, nullptr, -1);
        manifold_flush(cs, m);
    }
}

// TestTriangleAgainstAabb2 (LinearMath/btAabbUtil2.h)
RL_HDI bool tri_vs_aabb(const Tri& t, V3 mn, V3 mx) {
    if (fminf_(fminf_(t.v0.x, t.v1.x), t.v2.x) > mx.x) return false;
    if (fmaxf_(fmaxf_(t.v0.x, t.v1.x), t.v2.x) < mn.x) return false;
    if (fminf_(fminf_(t.v0.z, t.v1.z), t.v2.z) > mx.z) return false;
    if (fmaxf_(fmaxf_(t.v0.z, t.v1.z), t.v2.z) < mn.z) return false;
    if (fminf_(fminf_(t.v0.y, t.v1.y), t.v2.y) > mx.y) return false;
    if (fmaxf_(fmaxf_(t.v0.y, t.v1.y), t.v2.y) < mn.y) return false;
    return true;
}

// SphereTriangleDetector::collide (SphereTriangleDetector.cpp:139-243, RocketSim-modified)
RL_HD inline bool sphere_triangle(V3 center, float radius, const Tri& t, float breaking, V3& point, V3& resultNormal, float& depth) {
    float radiusWithThreshold = radius + breaking;
    V3 normal = cross(t.v1 - t.v0, t.v2 - t.v0);
    float l2 = len2(normal);
    bool hasContact = false;
    V3 contactPoint;
    if (l2 >= kEps * kEps) {
        normal = normal / sqrtf(l2);
        V3 p1ToCentre = center - t.v0;
        float distanceFromPlane = dot(p1ToCentre, normal);
        if (distanceFromPlane < 0.f) { distanceFromPlane *= -1.f; normal = normal * -1.f; }
        if (distanceFromPlane < radiusWithThreshold) {
            // pointInTriangle (barycentric)
            V3 u = t.v1 - t.v0, v = t.v2 - t.v0;
            V3 n = cross(u, v);
            float nLenSq = dot(n, n);
            V3 w = center - t.v0;
            float gamma = dot(cross(u, w), n) / nLenSq;
            float beta = dot(cross(w, v), n) / nLenSq;
            float alpha = 1 - gamma - beta;
            bool inside = (0 <= alpha) && (alpha <= 1) && (0 <= beta) && (beta <= 1) && (0 <= gamma) && (gamma <= 1);
            if (inside) {
                hasContact = true;
                contactPoint = center - normal * distanceFromPlane;
            } else {
                float minDistSqr = radiusWithThreshold * radiusWithThreshold;
                V3 nearest = closest_pt_triangle(center, t.v0, t.v1, t.v2);
                float d2 = len2(nearest - center);
                if (d2 < minDistSqr) { hasContact = true; contactPoint = nearest; }
            }
        }
    }
    if (hasContact) {
        V3 contactToCentre = center - contactPoint;
        float distanceSqr = len2(contactToCentre);
        if (distanceSqr < radiusWithThreshold * radiusWithThreshold) {
            if (distanceSqr > kEps) {
                float distance = sqrtf(distanceSqr);
                resultNormal = normalized(contactToCentre);
                point = contactPoint;
                depth = -(radius - distance);
            } else {
                resultNormal = normal; point = contactPoint; depth = -radius;
            }
            return true;
        }
    }
    return false;
}

// support-plane early out of btConvexTriangleCallback::processTriangle (btConvexConcaveCollisionAlgorithm.cpp:100-138)
template <class Support>
RL_HDI bool tri_early_out(const Tri& t, float threshold, Support sup) {
    V3 n = normalized(cross(t.v1 - t.v0, t.v2 - t.v0));
    float dist = dot(n, t.v0) - dot(n, sup(n));
    if (dist > threshold) return true;
    n = n * -1.f;
    dist = dot(n, t.v0) - dot(n, sup(n));
    return dist > threshold;
}

// ball vs every mesh: btConvexConcaveCollisionAlgorithm + btSphereTriangleCollisionAlgorithm
RL_HD inline void sphere_meshes(CollideCtx& x, ContactSink& cs, const MeshSet& ms, V3 center, float radius, float breaking) {
    float am = radius + 0.08f;  // btSphereShape::getAabb
    V3 mn = center - V3(am, am, am), mx = center + V3(am, am, am);
    if (inside_free_box(ms, mn, mx)) return;
    auto leaf = [&](Manifold& m, const BvhNode& nd) {  // one leaf whose box overlaps the sphere's
        const Tri& t = ms.tris[nd.tri];
        if (!tri_vs_aabb(t, mn, mx)) return;
        auto sup = [&](V3 d) {  // btSphereShape::localGetSupportingVertex
            float l2 = len2(d);
            V3 dn = l2 < kEps * kEps ? V3(-1, -1, -1) : d;
            return center + normalized(dn) * radius;
        };
        if (tri_early_out(t, breaking, sup)) return;
        V3 point, normal; float depth;
        if (sphere_triangle(center, radius, t, breaking, point, normal, depth)) manifold_add(x, m, normal, point, depth, &ms, nd.tri);
    };
    int first, count;
    if (grid_lookup(ms, mn, mx, first, count)) {  // flat scan of the cell's leaf list: same leaves, same order as the walk below
        int j = 0;
        while (j < count) {
            const int mi = ms.gridList[first + j] >> 24;
            Manifold m; m.a = 0; m.b = -1; m.n = 0; m.breaking = breaking;
            for (; j < count && (ms.gridList[first + j] >> 24) == mi; j++) {
                const BvhNode& nd = ms.nodes[ms.gridList[first + j] & 0xffffff];
                if (aabb_overlap(nd.mn, nd.mx, mn, mx)) leaf(m, nd);
            }
            manifold_flush(cs, m);
        }
        return;
    }
    for (int mi = 0; mi < ms.numMeshes; mi++) {
        const BvhNode& root = ms.nodes[ms.nodeStart[mi]];
        if (!aabb_overlap(root.mn, root.mx, mn, mx)) continue;
        Manifold m; m.a = 0; m.b = -1; m.n = 0; m.breaking = breaking;
        for (int h = ms.hdrStart[mi]; h < ms.hdrStart[mi + 1]; h++) {
            int i = ms.hdrRoot[h], end = ms.hdrRoot[h] + ms.hdrSize[h];
            while (i < end) {
                const BvhNode& nd = ms.nodes[i];
                bool ov = aabb_overlap(nd.mn, nd.mx, mn, mx);
                if (nd.tri >= 0) {
                    if (ov) leaf(m, nd);
                    i++;
                } else {
                    i += ov ? 1 : nd.escape;
                }
            }
        }
        manifold_flush(cs, m);
    }
}

// car hitbox vs every mesh: btCompoundCollisionAlgorithm -> btConvexConcaveCollisionAlgorithm -> GJK per triangle
// one candidate triangle of the hitbox-vs-mesh narrowphase (shared by the direct walk and the candidate-list path)
// The geometric part is a pure function of (car pose, triangle) — the role kernel evaluates it for many (car, triangle)
// pairs at once, one pair per lane (engine.cu box_meshes_warp); the manifold bookkeeping stays with the car's own lane.
RL_HDI bool box_mesh_item(const CarS& c, const CarConsts& k, V3 boxCenter, V3 mn, V3 mx, const Tri& t, float breaking, const EpaCtx* ws, V3& normal,
                          V3& pointOnB, float& dist) {
    if (!tri_vs_aabb(t, mn, mx)) return false;
    auto sup = [&](V3 d) {  // btBoxShape::localGetSupportingVertex (with margin)
        V3 dl = tmul(d, c.rot);
        V3 v(dl.x >= 0 ? k.halfExt.x : -k.halfExt.x, dl.y >= 0 ? k.halfExt.y : -k.halfExt.y, dl.z >= 0 ? k.halfExt.z : -k.halfExt.z);
        return boxCenter + c.rot * v;
    };
    if (tri_early_out(t, breaking, sup)) return false;
    return box_triangle_contact(boxCenter, c.rot, k.coreHalf, k.boxMargin, t, breaking, ws, normal, pointOnB, dist);
}
RL_HDI void box_mesh_triangle(CollideCtx& x, Manifold& m, const MeshSet& ms, const CarS& c, const CarConsts& k, V3 boxCenter, V3 mn, V3 mx,
                              int triIdx, float breaking) {
    V3 normal, pointOnB; float dist;
    if (box_mesh_item(c, k, boxCenter, mn, mx, ms.tris[triIdx], breaking, x.epa, normal, pointOnB, dist))
        manifold_add(x, m, normal, pointOnB, dist, &ms, triIdx);
}

// hitbox vs the candidate list collected in Car::_PreTickUpdate (same leaves, same order as the direct walk below)
RL_HD inline void box_meshes_candidates(CollideCtx& x, ContactSink& cs, const MeshSet& ms, const MeshCands& cands, int ci, float breaking) {
    const CarS& c = x.a->cars[ci];
    const CarConsts& k = *x.k;
    V3 boxCenter = c.pos + c.rot * k.hitboxOffset;
    V3 ext(dot(vabs(c.rot.r[0]), k.halfExt), dot(vabs(c.rot.r[1]), k.halfExt), dot(vabs(c.rot.r[2]), k.halfExt));
    V3 mn = boxCenter - ext, mx = boxCenter + ext;
    int j = 0;
    while (j < cands.n) {
        const int mi = cands.node[j] >> 24;
        Manifold m; m.a = 1 + ci; m.b = -1; m.n = 0; m.breaking = breaking;
        for (; j < cands.n && (cands.node[j] >> 24) == mi; j++) {
            const BvhNode& nd = ms.nodes[cands.node[j] & 0xffffff];
            if (aabb_overlap(nd.mn, nd.mx, mn, mx)) box_mesh_triangle(x, m, ms, c, k, boxCenter, mn, mx, nd.tri, breaking);
        }
        manifold_flush(cs, m);
    }
}

RL_HD inline void box_meshes(CollideCtx& x, ContactSink& cs, const MeshSet& ms, int ci, float breaking) {
    const CarS& c = x.a->cars[ci];
    const CarConsts& k = *x.k;
    V3 boxCenter = c.pos + c.rot * k.hitboxOffset;
    // btBoxShape::getAabb -> btTransformAabb(halfExtentsWithoutMargin, margin, t)
    V3 ext(dot(vabs(c.rot.r[0]), k.halfExt), dot(vabs(c.rot.r[1]), k.halfExt), dot(vabs(c.rot.r[2]), k.halfExt));
    V3 mn = boxCenter - ext, mx = boxCenter + ext;
    if (inside_free_box(ms, mn, mx)) return;
    for (int mi = 0; mi < ms.numMeshes; mi++) {
        const BvhNode& root = ms.nodes[ms.nodeStart[mi]];
        if (!aabb_overlap(root.mn, root.mx, mn, mx)) continue;
        Manifold m; m.a = 1 + ci; m.b = -1; m.n = 0; m.breaking = breaking;
        for (int h = ms.hdrStart[mi]; h < ms.hdrStart[mi + 1]; h++) {
            int i = ms.hdrRoot[h], end = ms.hdrRoot[h] + ms.hdrSize[h];
            while (i < end) {
                const BvhNode& nd = ms.nodes[i];
                bool ov = aabb_overlap(nd.mn, nd.mx, mn, mx);
                if (nd.tri >= 0) {
                    if (ov) box_mesh_triangle(x, m, ms, c, k, boxCenter, mn, mx, nd.tri, breaking);
                    i++;
                } else {
                    i += ov ? 1 : nd.escape;
                }
            }
        }
        manifold_flush(cs, m);
    }
}

// ball vs car hitbox: GJK of (box core + 0.04) against (point + radius) == closest point on the core
RL_HD inline void car_ball(CollideCtx& x, ContactSink& cs, int ci, float breaking) {
    const CarS& c = x.a->cars[ci];
    const CarConsts& k = *x.k;
    float radius = x.cfg->mut.ballRadius * UU2BT;
    V3 boxCenter = c.pos + c.rot * k.hitboxOffset;
    V3 normal, pointOnB; float dist;
    if (box_sphere_contact(boxCenter, c.rot, k.coreHalf, k.boxMargin, x.ballPos, radius, breaking, x.epa, normal, pointOnB, dist)) {
        Manifold m; m.a = 1 + ci; m.b = 0; m.n = 0; m.breaking = breaking;
        manifold_add(x, m, normal, pointOnB, dist, nullptr, -1);
        manifold_flush(cs, m);
    }
}

// car c1 vs car c2 (c1 < c2).  Body order of the manifold: the pair reaches btCompoundCompoundCollisionAlgorithm as
// (c1, c2), which — the compounds have no dynamic AABB tree (btCompoundShape(false, 1), Car.cpp _BulletSetup) — defers to
// btCompoundCollisionAlgorithm: child box of c1 vs compound c2 goes through the SWAPPED compound algorithm, whose child
// call is (box of c2, box of c1), and the box-box algorithm creates the manifold for its own (body0, body1).  So body A
// of a car-car manifold (and box 1 of btBoxBoxDetector, and "car1" of the first bump test) is the HIGHER car.
RL_HD RL_NOINLINE inline void car_car(CollideCtx& x, ContactSink& cs, int c1, int c2, float breaking) {
    const CarS& A = x.a->cars[c2];
    const CarS& B = x.a->cars[c1];
    const CarConsts& k = *x.k;
    V3 ca = A.pos + A.rot * k.hitboxOffset, cb = B.pos + B.rot * k.hitboxOffset;
    BoxBoxResult r;
    box_box(ca, A.rot, k.halfExt, cb, B.rot, k.halfExt, r);
    if (r.n > 0) {
        Manifold m; m.a = 1 + c2; m.b = 1 + c1; m.n = 0; m.breaking = breaking;
        for (int i = 0; i < r.n; i++) manifold_add(x, m, r.normal, r.point[i], r.depth[i], nullptr, -1);
        manifold_flush(cs, m);
    }
}

}  // namespace rl

#include "rl_boxbox.h"
