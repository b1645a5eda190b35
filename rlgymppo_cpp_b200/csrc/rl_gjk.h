// rl_gjk.h — convex-convex closest points for the hitbox: box vs triangle, box vs ball, box vs box.
//
// The reference routes box-triangle and box-sphere through btConvexConvexAlgorithm ->
// btGjkPairDetector (B/BulletCollision/NarrowPhaseCollision/btGjkPairDetector.cpp:690-1000)
// with btVoronoiSimplexSolver, and box-box through btBoxBoxDetector.  This file restates
// those algorithms for the three concrete support mappings involved (box core, triangle,
// point), without any polymorphism: GJK distance on the margin-less cores, margins added
// afterwards; when the cores overlap or nearly touch, the reference's penetration-depth solver
// (GJK + EPA on the shapes with margins, rl_epa.h) takes over exactly as in btGjkPairDetector.
#pragma once
#include "rl_mesh.h"
#include "rl_epa.h"

namespace rl {

// closest point on triangle to p (embree variant the reference uses, SphereTriangleDetector.cpp:83-124)
RL_HD inline V3 closest_pt_triangle(V3 p, V3 a, V3 b, V3 c) {
    V3 ab = b - a, ac = c - a, ap = p - a;
    float d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0.f && d2 <= 0.f) return a;
    V3 bp = p - b;
    float d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0.f && d4 <= d3) return b;
    V3 cp = p - c;
    float d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0.f && d5 <= d6) return c;
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { float v = d1 / (d1 - d3); return a + ab * v; }
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { float v = d2 / (d2 - d6); return a + ac * v; }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) { float v = (d4 - d3) / ((d4 - d3) + (d5 - d6)); return b + (c - b) * v; }
    float denom = 1.f / (va + vb + vc);
    float v = vb * denom, w = vc * denom;
    return a + ab * v + ac * w;
}

// ---- Voronoi simplex (btVoronoiSimplexSolver) ------------------------------------------------
struct Simplex {
    V3 W[4], P[4], Q[4];
    int n;
    V3 lastW;
    V3 cachedV, cachedP1, cachedP2;
    float bc[4];
    bool valid;
    // vertices are uninitialised (see NoInit; slots < n are always written first); the cached results start at zero
    RL_HDI Simplex() : W{V3(NoInit()), V3(NoInit()), V3(NoInit()), V3(NoInit())}, P{V3(NoInit()), V3(NoInit()), V3(NoInit()), V3(NoInit())},
                       Q{V3(NoInit()), V3(NoInit()), V3(NoInit()), V3(NoInit())}, n(0), bc{0, 0, 0, 0}, valid(false) {}
};

struct SubResult { float bc[4]; bool used[4]; V3 closest; bool degenerate; };

RL_HD inline void closest_pt_tri_origin(V3 a, V3 b, V3 c, SubResult& r) {
    for (int i = 0; i < 4; i++) { r.used[i] = false; r.bc[i] = 0.f; }
    V3 p(0, 0, 0);
    V3 ab = b - a, ac = c - a, ap = p - a;
    float d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0.f && d2 <= 0.f) { r.closest = a; r.used[0] = true; r.bc[0] = 1; return; }
    V3 bp = p - b;
    float d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0.f && d4 <= d3) { r.closest = b; r.used[1] = true; r.bc[1] = 1; return; }
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
        float v = d1 / (d1 - d3);
        r.closest = a + ab * v; r.used[0] = r.used[1] = true; r.bc[0] = 1 - v; r.bc[1] = v; return;
    }
    V3 cp = p - c;
    float d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0.f && d5 <= d6) { r.closest = c; r.used[2] = true; r.bc[2] = 1; return; }
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
        float w = d2 / (d2 - d6);
        r.closest = a + ac * w; r.used[0] = r.used[2] = true; r.bc[0] = 1 - w; r.bc[2] = w; return;
    }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        r.closest = b + (c - b) * w; r.used[1] = r.used[2] = true; r.bc[1] = 1 - w; r.bc[2] = w; return;
    }
    float denom = 1.f / (va + vb + vc);
    float v = vb * denom, w = vc * denom;
    r.closest = a + ab * v + ac * w;
    r.used[0] = r.used[1] = r.used[2] = true;
    r.bc[0] = 1 - v - w; r.bc[1] = v; r.bc[2] = w;
}

RL_HDI int point_outside_plane(V3 p, V3 a, V3 b, V3 c, V3 d) {
    V3 normal = cross(b - a, c - a);
    float signp = dot(p - a, normal), signd = dot(d - a, normal);
    if (signd * signd < (1e-4f * 1e-4f)) return -1;
    return signp * signd < 0.f;
}

RL_HD inline bool closest_pt_tetra_origin(V3 a, V3 b, V3 c, V3 d, SubResult& f) {
    V3 p(0, 0, 0);
    f.closest = p; f.degenerate = false;
    for (int i = 0; i < 4; i++) { f.used[i] = true; f.bc[i] = 0; }
    int oABC = point_outside_plane(p, a, b, c, d), oACD = point_outside_plane(p, a, c, d, b);
    int oADB = point_outside_plane(p, a, d, b, c), oBDC = point_outside_plane(p, b, d, c, a);
    if (oABC < 0 || oACD < 0 || oADB < 0 || oBDC < 0) { f.degenerate = true; return false; }
    if (!oABC && !oACD && !oADB && !oBDC) return false;
    float best = 3.402823466e+38f;
    SubResult t;
    if (oABC) {
        closest_pt_tri_origin(a, b, c, t);
        float sq = dot(t.closest, t.closest);
        if (sq < best) { best = sq; f.closest = t.closest; f.used[0] = t.used[0]; f.used[1] = t.used[1]; f.used[2] = t.used[2]; f.used[3] = false;
            f.bc[0] = t.bc[0]; f.bc[1] = t.bc[1]; f.bc[2] = t.bc[2]; f.bc[3] = 0; }
    }
    if (oACD) {
        closest_pt_tri_origin(a, c, d, t);
        float sq = dot(t.closest, t.closest);
        if (sq < best) { best = sq; f.closest = t.closest; f.used[0] = t.used[0]; f.used[1] = false; f.used[2] = t.used[1]; f.used[3] = t.used[2];
            f.bc[0] = t.bc[0]; f.bc[1] = 0; f.bc[2] = t.bc[1]; f.bc[3] = t.bc[2]; }
    }
    if (oADB) {
        closest_pt_tri_origin(a, d, b, t);
        float sq = dot(t.closest, t.closest);
        if (sq < best) { best = sq; f.closest = t.closest; f.used[0] = t.used[0]; f.used[1] = t.used[2]; f.used[2] = false; f.used[3] = t.used[1];
            f.bc[0] = t.bc[0]; f.bc[1] = t.bc[2]; f.bc[2] = 0; f.bc[3] = t.bc[1]; }
    }
    if (oBDC) {
        closest_pt_tri_origin(b, d, c, t);
        float sq = dot(t.closest, t.closest);
        if (sq < best) { best = sq; f.closest = t.closest; f.used[0] = false; f.used[1] = t.used[0]; f.used[2] = t.used[2]; f.used[3] = t.used[1];
            f.bc[0] = 0; f.bc[1] = t.bc[0]; f.bc[2] = t.bc[2]; f.bc[3] = t.bc[1]; }
    }
    return true;
}

RL_HDI void simplex_remove(Simplex& s, int i) { s.n--; s.W[i] = s.W[s.n]; s.P[i] = s.P[s.n]; s.Q[i] = s.Q[s.n]; }
RL_HDI void simplex_reduce(Simplex& s, const bool* used) {
    if (s.n >= 4 && !used[3]) simplex_remove(s, 3);
    if (s.n >= 3 && !used[2]) simplex_remove(s, 2);
    if (s.n >= 2 && !used[1]) simplex_remove(s, 1);
    if (s.n >= 1 && !used[0]) simplex_remove(s, 0);
}

// btVoronoiSimplexSolver::updateClosestVectorAndPoints
RL_HD inline bool simplex_closest(Simplex& s, V3& v) {
    bool valid = false;
    switch (s.n) {
    case 0: valid = false; break;
    case 1:
        s.cachedP1 = s.P[0]; s.cachedP2 = s.Q[0]; s.cachedV = s.cachedP1 - s.cachedP2;
        valid = true; break;
    case 2: {
        V3 from = s.W[0], to = s.W[1];
        V3 diff = V3() - from, vv = to - from;
        float t = dot(vv, diff);
        bool used[4] = {false, false, false, false};
        if (t > 0) {
            float dotVV = dot(vv, vv);
            if (t < dotVV) { t /= dotVV; used[0] = used[1] = true; }
            else { t = 1; used[1] = true; }
        } else { t = 0; used[0] = true; }
        s.cachedP1 = s.P[0] + (s.P[1] - s.P[0]) * t;
        s.cachedP2 = s.Q[0] + (s.Q[1] - s.Q[0]) * t;
        s.cachedV = s.cachedP1 - s.cachedP2;
        simplex_reduce(s, used);
        valid = (1 - t) >= 0.f && t >= 0.f;
        break;
    }
    case 3: {
        SubResult r; r.degenerate = false;
        closest_pt_tri_origin(s.W[0], s.W[1], s.W[2], r);
        s.cachedP1 = s.P[0] * r.bc[0] + s.P[1] * r.bc[1] + s.P[2] * r.bc[2];
        s.cachedP2 = s.Q[0] * r.bc[0] + s.Q[1] * r.bc[1] + s.Q[2] * r.bc[2];
        s.cachedV = s.cachedP1 - s.cachedP2;
        simplex_reduce(s, r.used);
        valid = r.bc[0] >= 0.f && r.bc[1] >= 0.f && r.bc[2] >= 0.f && r.bc[3] >= 0.f;
        break;
    }
    case 4: {
        SubResult r;
        bool sep = closest_pt_tetra_origin(s.W[0], s.W[1], s.W[2], s.W[3], r);
        if (sep) {
            s.cachedP1 = s.P[0] * r.bc[0] + s.P[1] * r.bc[1] + s.P[2] * r.bc[2] + s.P[3] * r.bc[3];
            s.cachedP2 = s.Q[0] * r.bc[0] + s.Q[1] * r.bc[1] + s.Q[2] * r.bc[2] + s.Q[3] * r.bc[3];
            s.cachedV = s.cachedP1 - s.cachedP2;
            simplex_reduce(s, r.used);
            valid = r.bc[0] >= 0.f && r.bc[1] >= 0.f && r.bc[2] >= 0.f && r.bc[3] >= 0.f;
        } else {
            if (r.degenerate) valid = false;
            else { valid = true; s.cachedV = V3(0, 0, 0); }
        }
        break;
    }
    default: valid = false;
    }
    s.valid = valid;
    v = s.cachedV;
    return valid;
}

RL_HDI bool simplex_contains(const Simplex& s, V3 w) {
    for (int i = 0; i < s.n; i++)
        if (s.W[i].x == w.x && s.W[i].y == w.y && s.W[i].z == w.z) return true;
    return w.x == s.lastW.x && w.y == s.lastW.y && w.z == s.lastW.z;
}

// The front half of btGjkPairDetector::getClosestPointsNonVirtual (btGjkPairDetector.cpp:690-850) for A = box core and a convex B
// given by a support functor in world space: GJK distance between the margin-less cores in the frame shifted by `offset`.
struct GjkOut {
    bool valid;        // isValid: a usable separating axis was found
    int degenerate;    // m_degenerateSimplex
    V3 axis;           // m_cachedSeparatingAxis (unnormalised)
    V3 pB;             // closest point on B's core, shifted frame
    float sqDist;      // squaredDistance
};
template <class SupportB>
RL_HD inline void gjk_box_core(V3 cA, const M3& rot, V3 coreHalf, V3 offset, SupportB supB, float maxDistSq, GjkOut& o) {
    Simplex s; s.n = 0; s.lastW = V3(1e18f, 1e18f, 1e18f);
    V3 sepAxis(0, 1, 0);
    float squaredDistance = 1e18f;
    bool checkSimplex = false;
    int degenerate = 0;
    const float REL_ERROR2 = 1.0e-6f;
    for (int iter = 0; iter <= 1000; iter++) {
        V3 dirA = tmul(-sepAxis, rot);
        V3 pInA(dirA.x >= 0 ? coreHalf.x : -coreHalf.x, dirA.y >= 0 ? coreHalf.y : -coreHalf.y, dirA.z >= 0 ? coreHalf.z : -coreHalf.z);
        V3 pWorld = cA + rot * pInA;
        V3 qWorld = supB(sepAxis) - offset;
        V3 w = pWorld - qWorld;
        float delta = dot(sepAxis, w);
        if (delta > 0.f && delta * delta > squaredDistance * maxDistSq) { degenerate = 10; checkSimplex = true; break; }
        if (simplex_contains(s, w)) { degenerate = 1; checkSimplex = true; break; }
        float f0 = squaredDistance - delta, f1 = squaredDistance * REL_ERROR2;
        if (f0 <= f1) { degenerate = f0 <= 0.f ? 2 : 11; checkSimplex = true; break; }
        s.lastW = w; s.W[s.n] = w; s.P[s.n] = pWorld; s.Q[s.n] = qWorld; s.n++;
        V3 newAxis;
        if (!simplex_closest(s, newAxis)) { degenerate = 3; checkSimplex = true; break; }
        if (len2(newAxis) < REL_ERROR2) { sepAxis = newAxis; degenerate = 6; checkSimplex = true; break; }
        float prev = squaredDistance;
        squaredDistance = len2(newAxis);
        if (prev - squaredDistance <= kEps * prev) { checkSimplex = true; degenerate = 12; break; }
        sepAxis = newAxis;
        if (s.n == 4) { degenerate = 13; break; }
    }
    o.valid = false; o.degenerate = degenerate; o.axis = sepAxis; o.sqDist = squaredDistance; o.pB = V3();
    if (checkSimplex) {
        V3 dummy; simplex_closest(s, dummy);
        o.pB = s.cachedP2;
        if (len2(sepAxis) > kEps * kEps) o.valid = true;
    }
}

// Shape B of the penetration-depth search in its own frame (identity basis): a triangle (margin 0) or the ball (point core,
// margin = radius) — btConvexShape::localGetSupportVertex[WithoutMargin]NonVirtual for the two.  One non-template description,
// so the search below exists once in the kernel image.
struct ConvexB {
    V3 v0, v1, v2;
    float radius;
    int32_t sphere;
    RL_HDI V3 operator()(V3 dir, bool withMargin) const {
        V3 dn = dir;
        if (withMargin) {
            if (len2(dn) < kEps * kEps) dn = V3(-1, -1, -1);
            dn = normalized(dn);
        }
        if (sphere) return withMargin ? dn * radius : V3(0, 0, 0);
        float d0 = dot(dn, v0), d1 = dot(dn, v1), d2 = dot(dn, v2);
        int mi = d0 < d1 ? (d1 < d2 ? 2 : 1) : (d0 < d2 ? 2 : 0);
        return mi == 0 ? v0 : (mi == 1 ? v1 : v2);
    }
};

#if defined(RL_DEBUG_CONTACTS) && !defined(__CUDA_ARCH__)
static long g_dbg_pen_kind[2] = {0, 0};  // host debugging: calls with overlapping cores / with a degenerate-small GJK distance
#endif
// btGjkPairDetector.cpp:860-940: the penetration-depth solver and the "only replace when deeper / closer" rules.  Rare and
// long: out of line, one copy.  Updates (isValid, normalInB, pointOnB, distance) in place.
RL_HD RL_NOINLINE inline void pair_penetration(const EpaCtx* ws, const M3& rot, V3 cA, V3 oB, V3 coreHalf, float marginA, float marginB, const ConvexB& shapeB,
                                               bool& isValid, V3& normalInB, V3& pointOnB, float& distance) {
    const float margin = marginA + marginB;
#if defined(RLG_EPA_TIMING) && defined(__CUDA_ARCH__)
    const long long epaT0 = clock64();
#endif
    with_epa_ws(ws, [&](const EpaWs& w) {
        Mink<ConvexB> sh(rot, cA, oB, coreHalf, marginA, shapeB);
        PenResult r;
        const int pr = calc_pen_depth(w, sh, cA, oB, r);
        if (pr == -2) return false;
        const V3 v = r.normal;  // m_cachedSeparatingAxis after calcPenDepth (A's local frame: reference quirk)
        if (pr == 1) {
            V3 tmpN = r.witnessB - r.witnessA;
            float lenSqr = len2(tmpN);
            if (lenSqr <= kEps * kEps) { tmpN = v; lenSqr = len2(v); }
            if (lenSqr > kEps * kEps) {
                tmpN = tmpN / sqrtf(lenSqr);
                const float distance2 = -len(r.witnessA - r.witnessB);
                if (!isValid || distance2 < distance) { distance = distance2; pointOnB = r.witnessB; normalInB = tmpN; isValid = true; }
            }
        } else if (len2(v) > 0.f) {
            const float distance2 = len(r.witnessA - r.witnessB) - margin;
            if (!isValid || distance2 < distance) {
                distance = distance2;
                pointOnB = r.witnessB + v * marginB;
                normalInB = normalized(v);
                isValid = true;
            }
        }
        return true;
    });
#if defined(RLG_EPA_TIMING) && defined(__CUDA_ARCH__)
    atomicAdd(&g_epa_timing[0], (unsigned long long)(clock64() - epaT0));
    atomicAdd(&g_epa_timing[1], 1ULL);
#endif
}

// The back half (btGjkPairDetector.cpp:850-1000): penetration-depth solver when the cores overlap or the distance is
// degenerate-small, the contact-normal direction fix, the distance gate.  In: the GJK result as (valid, normal, pointOnB in
// the shifted frame, distance); oB = B's origin in the shifted frame, posB = centre of B's bounding box.
RL_HDI bool pair_finish(bool isValid, bool degenerate, V3 normalInB, V3 pointOnB, float distance, V3 cA, const M3& rot, V3 coreHalf, float marginA,
                        float marginB, V3 oB, const ConvexB& shapeB, V3 posB, V3 offset, float maxDistSq, const EpaCtx* ws, V3& outNormal, V3& outPoint,
                        float& outDist) {
    const bool catchDegenerate = degenerate && (distance + (marginA + marginB)) < 0.01f;
#if defined(RL_DEBUG_CONTACTS) && !defined(__CUDA_ARCH__)
    if (!isValid) g_dbg_epa_hist[0 + 0] += 0, g_dbg_pen_kind[0]++; else if (catchDegenerate) g_dbg_pen_kind[1]++;
#endif
    if (!isValid || catchDegenerate) pair_penetration(ws, rot, cA, oB, coreHalf, marginA, marginB, shapeB, isValid, normalInB, pointOnB, distance);
    if (isValid && (distance < 0 || distance * distance < maxDistSq)) {
        // m_fixContactNormalDirection: the normal must point from B's bounding-box centre towards A's
        if (dot(cA - posB, normalInB) < 0.f) normalInB = normalInB * -1.f;
        outNormal = normalInB; outPoint = pointOnB + offset; outDist = distance;
        return true;
    }
    return false;
}

// ---- box vs sphere (car hitbox vs ball) ---------------------------------------------------------
// A = box (margin 0.04), B = sphere (point core, margin = radius).  GJK between a box core and a point converges to the
// closest point on the core; written in closed form.  Centre within 0.01 of the core (or inside it): penetration solver.
RL_HD RL_NOINLINE inline bool box_sphere_contact(V3 boxCenter, const M3& rot, V3 core, float marginA, V3 sphereCenter, float radius, float breaking,
                                                 const EpaCtx* ws, V3& normalOnB, V3& pointOnB, float& dist) {
    V3 l = tmul(sphereCenter - boxCenter, rot);
    V3 q(clampf(l.x, -core.x, core.x), clampf(l.y, -core.y, core.y), clampf(l.z, -core.z, core.z));
    V3 d = q - l;  // from sphere centre (B) to box core (A)
    float d2 = len2(d);
    float margin = marginA + radius;
    float maxDist = margin + breaking;
    if (d2 > maxDist * maxDist) return false;
    const float dl = sqrtf(d2);
    if (d2 > kEps * kEps && dl >= 0.01f) {
        V3 nl = d * (1.f / dl);
        normalOnB = rot * nl;
        pointOnB = sphereCenter + normalOnB * radius;
        dist = dl - margin;
        return true;
    }
    const V3 offset = (boxCenter + sphereCenter) * 0.5f;
    const bool valid = d2 > kEps * kEps;
    V3 n0 = valid ? rot * (d * (1.f / dl)) : V3();
    ConvexB sb; sb.sphere = 1; sb.radius = radius;
    return pair_finish(valid, true, n0, (sphereCenter - offset) + n0 * radius, dl - margin, boxCenter - offset, rot, core, marginA, radius, sphereCenter - offset,
                       sb, sphereCenter - offset, offset, maxDist * maxDist, ws, normalOnB, pointOnB, dist);
}

// ---- box vs triangle -------------------------------------------------------------------------------
RL_HD RL_NOINLINE inline bool box_triangle_contact(V3 boxCenter, const M3& rot, V3 core, float marginA, const Tri& t, float breaking, const EpaCtx* ws,
                                                   V3& normalOnB, V3& pointOnB, float& dist) {
    const float maxDist = marginA + 0.f + breaking;
    const V3 v0 = t.v0, v1 = t.v1, v2 = t.v2;
    auto supWorld = [&](V3 axis) {  // btTriangleShape::localGetSupportingVertexWithoutMargin(axis * basisB), basisB = I
        float d0 = dot(axis, v0), d1 = dot(axis, v1), d2 = dot(axis, v2);
        int mi = d0 < d1 ? (d1 < d2 ? 2 : 1) : (d0 < d2 ? 2 : 0);
        return mi == 0 ? v0 : (mi == 1 ? v1 : v2);
    };
    const V3 offset = boxCenter * 0.5f;  // (originA + originB) / 2, originB = 0
    const V3 cA = boxCenter - offset;
    GjkOut g;
    gjk_box_core(cA, rot, core, offset, supWorld, maxDist * maxDist, g);
    bool isValid = false;
    V3 normalInB(0, 0, 0), pB = g.pB;
    float distance = 0.f;
    if (g.valid) {
        const float rlen = 1.f / sqrtf(len2(g.axis));
        normalInB = g.axis * rlen;
        const float s = sqrtf(g.sqDist);
        pB = g.pB + g.axis * (0.f / s);  // marginB = 0
        distance = (1.f / rlen) - marginA;
        isValid = true;
    }
    // centre of the triangle's bounding box in the shifted frame (btPolyhedralConvexShape::getAabb via the support mapping, margin 0)
    const V3 oB = V3() - offset;
    const V3 posB = ((vmin(vmin(v0, v1), v2) + oB) + (vmax(vmax(v0, v1), v2) + oB)) * 0.5f;
    ConvexB sb; sb.sphere = 0; sb.radius = 0.f; sb.v0 = v0; sb.v1 = v1; sb.v2 = v2;
    return pair_finish(isValid, g.degenerate != 0, normalInB, pB, distance, cA, rot, core, marginA, 0.f, oB, sb, posB, offset, maxDist * maxDist, ws, normalOnB,
                       pointOnB, dist);
}

// ---- box vs box (btBoxBoxDetector / ODE dBoxBox2) --------------------------------------------------
struct BoxBoxResult { int n; V3 normal; V3 point[4]; float depth[4]; };
RL_HD inline void box_box(V3 ca, const M3& ra, V3 ha, V3 cb, const M3& rb, V3 hb, BoxBoxResult& out);

// ---- internal edge adjustment ------------------------------------------------------------------------
struct Contact;
RL_HD inline void adjust_internal_edge(Contact& cp, const MeshSet& ms, int tri);

}  // namespace rl
