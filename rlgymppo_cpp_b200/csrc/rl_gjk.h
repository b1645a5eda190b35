// rl_gjk.h — convex-convex closest points for the hitbox: box vs triangle, box vs ball, box vs box.
//
// The reference routes box-triangle and box-sphere through btConvexConvexAlgorithm ->
// btGjkPairDetector (B/BulletCollision/NarrowPhaseCollision/btGjkPairDetector.cpp:690-1000)
// with btVoronoiSimplexSolver, and box-box through btBoxBoxDetector.  This file restates
// those algorithms for the three concrete support mappings involved (box core, triangle,
// point), without any polymorphism: GJK distance on the margin-less cores, margins added
// afterwards; when the cores overlap (the reference switches to EPA on the rounded box) a
// 13-axis separating-axis search on box vs triangle yields the minimum translation.
#pragma once
#include "rl_mesh.h"

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

// GJK between a box core (A) and a convex B given by a support functor in world space.
// Returns true with separated cores: pA/pB closest points on the cores, axis = pA - pB.
// Returns false if the cores overlap/touch (penetration path).
template <class SupportB>
RL_HD inline bool gjk_box_core(V3 boxCenter, const M3& rot, V3 coreHalf, V3 originB, SupportB supB, float maxDistSq, V3& pA, V3& pB, V3& axis, float& sqDist, bool& tooFar) {
    V3 offset = (boxCenter + originB) * 0.5f;
    V3 cA = boxCenter - offset;
    Simplex s; s.n = 0; s.lastW = V3(1e18f, 1e18f, 1e18f);
    V3 sepAxis(0, 1, 0);
    float squaredDistance = 1e18f;
    bool checkSimplex = false;
    int degenerate = 0;
    tooFar = false;
    const float REL_ERROR2 = 1.0e-6f;
    for (int iter = 0; iter < 64; iter++) {
        V3 dirA = tmul(-sepAxis, rot);
        V3 pInA(dirA.x >= 0 ? coreHalf.x : -coreHalf.x, dirA.y >= 0 ? coreHalf.y : -coreHalf.y, dirA.z >= 0 ? coreHalf.z : -coreHalf.z);
        V3 pWorld = cA + rot * pInA;
        V3 qWorld = supB(sepAxis) - offset;
        V3 w = pWorld - qWorld;
        float delta = dot(sepAxis, w);
        if (delta > 0.f && delta * delta > squaredDistance * maxDistSq) { degenerate = 10; checkSimplex = true; tooFar = true; break; }
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
    (void)degenerate;
    if (checkSimplex) {
        V3 dummy; simplex_closest(s, dummy);
        pA = s.cachedP1 + offset; pB = s.cachedP2 + offset;
        float lenSqr = len2(sepAxis);
        if (lenSqr > kEps * kEps) { axis = sepAxis; sqDist = squaredDistance; return true; }
    }
    return false;
}

// ---- box vs sphere (car hitbox vs ball) ---------------------------------------------------------
// A = box (margin 0.04), B = sphere (point core, margin = radius).  GJK between a box core and a point
// converges to the closest point on the core; written in closed form.
RL_HD RL_NOINLINE inline bool box_sphere_contact(V3 boxCenter, const M3& rot, V3 halfExt, V3 core, float marginA, V3 sphereCenter, float radius, float breaking,
                                     V3& normalOnB, V3& pointOnB, float& dist) {
    V3 l = tmul(sphereCenter - boxCenter, rot);
    V3 q(clampf(l.x, -core.x, core.x), clampf(l.y, -core.y, core.y), clampf(l.z, -core.z, core.z));
    V3 d = q - l;  // from sphere centre (B) to box core (A)
    float d2 = len2(d);
    float margin = marginA + radius;
    float maxDist = margin + breaking;
    if (d2 > maxDist * maxDist) return false;
    if (d2 > kEps * kEps) {
        float dl = sqrtf(d2);
        V3 nl = d * (1.f / dl);
        normalOnB = rot * nl;
        pointOnB = sphereCenter + normalOnB * radius;
        dist = dl - margin;
        return true;
    }
    // sphere centre inside the core: minimum translation through the nearest face (stands in for EPA)
    float best = 1e18f; int ax = 0; float sg = 1.f;
    for (int a = 0; a < 3; a++) {
        float dp = halfExt[a] - l[a], dn = halfExt[a] + l[a];
        if (dp < best) { best = dp; ax = a; sg = 1.f; }
        if (dn < best) { best = dn; ax = a; sg = -1.f; }
    }
    V3 nl(0, 0, 0); nl[ax] = -sg;  // from sphere towards the box interior
    normalOnB = rot * nl;
    pointOnB = sphereCenter + normalOnB * radius;
    dist = -(best + radius);
    return true;
}

// ---- box vs triangle -------------------------------------------------------------------------------
RL_HDI void project_box(V3 axis, V3 center, const M3& rot, V3 half, float& mn, float& mx) {
    float c = dot(axis, center);
    float r = fabsf(dot(axis, rot.col(0))) * half.x + fabsf(dot(axis, rot.col(1))) * half.y + fabsf(dot(axis, rot.col(2))) * half.z;
    mn = c - r; mx = c + r;
}

// cores overlap: 13-axis SAT between the full box and the triangle -> minimum translation (B -> A)
RL_HD inline bool box_triangle_sat(V3 boxCenter, const M3& rot, V3 halfExt, const Tri& t, V3& normalOnB, V3& pointOnB, float& dist) {
    V3 e[3] = {t.v1 - t.v0, t.v2 - t.v1, t.v0 - t.v2};
    V3 axes[13];
    int na = 0;
    axes[na++] = cross(e[0], t.v2 - t.v0);
    for (int i = 0; i < 3; i++) axes[na++] = rot.col(i);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) axes[na++] = cross(rot.col(i), e[j]);
    float bestDepth = 1e18f; V3 bestAxis(0, 0, 1); int bestIdx = -1;
    for (int i = 0; i < na; i++) {
        float l2 = len2(axes[i]);
        if (l2 < 1e-10f) continue;
        V3 ax = axes[i] * (1.f / sqrtf(l2));
        float bmn, bmx; project_box(ax, boxCenter, rot, halfExt, bmn, bmx);
        float p0 = dot(ax, t.v0), p1 = dot(ax, t.v1), p2 = dot(ax, t.v2);
        float tmn = fminf_(fminf_(p0, p1), p2), tmx = fmaxf_(fmaxf_(p0, p1), p2);
        // push A (box) along +ax out of B: depth = tmx - bmn ; along -ax: bmx - tmn
        float dPos = tmx - bmn, dNeg = bmx - tmn;
        if (dPos < 0.f || dNeg < 0.f) return false;  // separated
        float bias = i >= 4 ? 1.0001f : 1.f;        // prefer face axes on ties, like most SAT implementations
        if (dPos * bias < bestDepth) { bestDepth = dPos * bias; bestAxis = ax; bestIdx = i; }
        if (dNeg * bias < bestDepth) { bestDepth = dNeg * bias; bestAxis = -ax; bestIdx = i; }
    }
    if (bestIdx < 0) return false;
    normalOnB = bestAxis;
    // witness on the triangle: for box-face axes the deepest triangle vertex, otherwise the deepest box vertex projected
    float depth = bestDepth / (bestIdx >= 4 ? 1.0001f : 1.f);
    if (bestIdx >= 1 && bestIdx <= 3) {
        float p0 = dot(bestAxis, t.v0), p1 = dot(bestAxis, t.v1), p2 = dot(bestAxis, t.v2);
        float mx = fmaxf_(fmaxf_(p0, p1), p2);
        V3 sum(0, 0, 0); int cnt = 0;
        if (mx - p0 < 1e-4f) { sum += t.v0; cnt++; }
        if (mx - p1 < 1e-4f) { sum += t.v1; cnt++; }
        if (mx - p2 < 1e-4f) { sum += t.v2; cnt++; }
        pointOnB = sum * (1.f / (float)cnt);
    } else {
        V3 dl = tmul(-bestAxis, rot);
        V3 sum(0, 0, 0); int cnt = 0;
        for (int v = 0; v < 8; v++) {
            V3 lv((v & 1) ? halfExt.x : -halfExt.x, (v & 2) ? halfExt.y : -halfExt.y, (v & 4) ? halfExt.z : -halfExt.z);
            float sup = fabsf(dl.x) * halfExt.x + fabsf(dl.y) * halfExt.y + fabsf(dl.z) * halfExt.z;
            if (sup - dot(dl, lv) < 1e-4f) { sum += lv; cnt++; }
        }
        V3 pa = boxCenter + rot * (sum * (1.f / (float)cnt));
        pointOnB = closest_pt_triangle(pa + bestAxis * depth, t.v0, t.v1, t.v2);
    }
    dist = -depth;
    return true;
}

RL_HD RL_NOINLINE inline bool box_triangle_contact(V3 boxCenter, const M3& rot, V3 halfExt, V3 core, float marginA, const Tri& t, float breaking,
                                       V3& normalOnB, V3& pointOnB, float& dist) {
    float maxDist = marginA + 0.f + breaking;
    auto supB = [&](V3 axis) {  // btTriangleShape::localGetSupportingVertexWithoutMargin(axis * basisB), basisB = I
        float d0 = dot(axis, t.v0), d1 = dot(axis, t.v1), d2 = dot(axis, t.v2);
        int mi = d0 < d1 ? (d1 < d2 ? 2 : 1) : (d0 < d2 ? 2 : 0);
        return mi == 0 ? t.v0 : (mi == 1 ? t.v1 : t.v2);
    };
    V3 pA, pB, axis; float sq; bool tooFar;
    if (gjk_box_core(boxCenter, rot, core, V3(), supB, maxDist * maxDist, pA, pB, axis, sq, tooFar)) {
        float lenSqr = len2(axis);
        float rlen = 1.f / sqrtf(lenSqr);
        V3 n = axis * rlen;
        float s = sqrtf(sq);
        V3 pointB = pB + axis * (0.f / s);
        float distance = (1.f / rlen) - marginA;
        bool catchDegenerate = (distance + marginA) < 0.01f;  // m_catchDegeneracies path re-checks with the penetration solver
        if (!catchDegenerate) {
            if (distance < 0 || distance * distance < maxDist * maxDist) { normalOnB = n; pointOnB = pointB; dist = distance; return true; }
            return false;
        }
        V3 n2, p2; float d2;
        if (box_triangle_sat(boxCenter, rot, halfExt, t, n2, p2, d2) && d2 < distance) { normalOnB = n2; pointOnB = p2; dist = d2; return true; }
        normalOnB = n; pointOnB = pointB; dist = distance;
        return distance < 0 || distance * distance < maxDist * maxDist;
    }
    return box_triangle_sat(boxCenter, rot, halfExt, t, normalOnB, pointOnB, dist);
}

// ---- box vs box (btBoxBoxDetector / ODE dBoxBox2) --------------------------------------------------
struct BoxBoxResult { int n; V3 normal; V3 point[4]; float depth[4]; };
RL_HD inline void box_box(V3 ca, const M3& ra, V3 ha, V3 cb, const M3& rb, V3 hb, BoxBoxResult& out);

// ---- internal edge adjustment ------------------------------------------------------------------------
struct Contact;
RL_HD inline void adjust_internal_edge(Contact& cp, const MeshSet& ms, int tri);

}  // namespace rl
