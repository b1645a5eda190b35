// rl_boxbox.h — car-car hitbox contacts (btBoxBoxDetector) and internal-edge normal adjustment
// (btInternalEdgeUtility).  Included at the end of rl_collide.h.
#pragma once
#include "rl_collide.h"

namespace rl {

// ---- btAdjustInternalEdgeContacts (B/BulletCollision/CollisionDispatch/btInternalEdgeUtility.cpp:414-800) ----
// triFlags bits: 1/2/4 = edge v0v1 / v1v2 / v2v0 convex, 8/16/32 = swap normalB; triEdgeAngles default 2*pi.
RL_HDI V3 nearest_on_segment(V3 p, V3 l0, V3 l1) {
    V3 d = l1 - l0;
    if (len2(d) < kEps * kEps) return l0;
    float t = dot(p - l0, d) / dot(d, d);
    if (t < 0) t = 0; else if (t > 1) t = 1;
    return l0 + d * t;
}
RL_HDI V3 quat_rotate(Quat q, V3 v) {
    // q * v
    Quat t(q.w * v.x + q.y * v.z - q.z * v.y, q.w * v.y + q.z * v.x - q.x * v.z, q.w * v.z + q.x * v.y - q.y * v.x,
           -q.x * v.x - q.y * v.y - q.z * v.z);
    Quat inv(-q.x, -q.y, -q.z, q.w);
    Quat r = t * inv;
    return V3(r.x, r.y, r.z);
}
RL_HDI float bt_angle(V3 edgeA, V3 normalA, V3 normalB) { return rl_atan2(dot(normalB, edgeA), dot(normalB, normalA)); }

RL_HD inline bool clamp_normal(V3 edge, V3 triNormal, V3 localN, float correctedEdgeAngle, V3& out) {
    V3 edgeCross = normalized(cross(edge, triNormal));
    float curAngle = bt_angle(edgeCross, triNormal, localN);
    if (correctedEdgeAngle < 0) {
        if (curAngle < correctedEdgeAngle) {
            out = quat_to_mat(quat_axis_angle(edge, correctedEdgeAngle - curAngle)) * localN;
            return true;
        }
    }
    if (correctedEdgeAngle >= 0) {
        if (curAngle > correctedEdgeAngle) {
            out = quat_to_mat(quat_axis_angle(edge, correctedEdgeAngle - curAngle)) * localN;
            return true;
        }
    }
    return false;
}

RL_HD RL_NOINLINE inline void adjust_internal_edge(Contact& cp, const MeshSet& ms, int tri) {
    const float kTwoPi = 6.283185307179586232f;
    const float edgeDistanceThreshold = 0.1f;
    const Tri& t = ms.tris[tri];
    int flags = ms.triFlags[tri];
    const float* ang = ms.triEdgeAngles + tri * 3;
    V3 v[3] = {t.v0, t.v1, t.v2};
    V3 triNormal = normalized(cross(t.v1 - t.v0, t.v2 - t.v0));
    V3 contact = cp.posB;
    V3 localN = normalized(cp.normal);
    int bestedge = -1;
    float best = 1e18f;
    for (int e = 0; e < 3; e++) {
        if (fabsf(ang[e]) < kTwoPi) {
            float l = len(contact - nearest_on_segment(contact, v[e], v[(e + 1) % 3]));
            if (l < best) { bestedge = e; best = l; }
        }
    }
    bool isNearEdge = false;
    int numConcave = 0;
    for (int e = 0; e < 3; e++) {
        if (!(fabsf(ang[e]) < kTwoPi)) continue;
        V3 a = v[e], b = v[(e + 1) % 3];
        float l = len(contact - nearest_on_segment(contact, a, b));
        if (l < edgeDistanceThreshold && bestedge == e) {
            V3 edge = a - b;
            isNearEdge = true;
            if (ang[e] == 0.f) {
                numConcave++;
            } else {
                bool convex = (flags & (1 << e)) != 0;
                float swapFactor = convex ? 1.f : -1.f;
                V3 nA = triNormal * swapFactor;
                V3 computedNormalB = quat_rotate(quat_axis_angle(edge, ang[e]), triNormal);
                if (flags & (8 << e)) computedNormalB = computedNormalB * -1.f;
                V3 nB = computedNormalB * swapFactor;
                float NdotA = dot(localN, nA), NdotB = dot(localN, nB);
                bool backFacing = (NdotA < 0.f) && (NdotB < 0.f);
                if (backFacing) {
                    numConcave++;
                } else {
                    V3 ln = (e == 0) ? localN : cp.normal;  // edges 1/2 re-read the (un-normalised) world normal
                    V3 clamped;
                    if (clamp_normal(edge, triNormal * swapFactor, ln, ang[e], clamped)) {
                        if (dot(clamped, triNormal) > 0) {
                            cp.normal = clamped;
                            cp.posB = cp.posA - cp.normal * cp.dist;
                        }
                    }
                }
            }
        }
    }
    if (isNearEdge && numConcave > 0) {
        float d = dot(triNormal, localN);
        if (d < 0) return;
        cp.normal = triNormal;
        cp.posB = cp.posA - cp.normal * cp.dist;
    }
}

// ---- box vs box ------------------------------------------------------------------------------------------------------------------
// Contact manifold of two oriented boxes with the decisions of Bullet's btBoxBoxDetector
// (B/BulletCollision/CollisionDispatch/btBoxBoxDetector.cpp:260-720) so that car-car contacts reproduce the reference:
//   1. separating-axis search over the 3 + 3 face normals and the 9 edge-pair cross products, in that order; an edge axis only
//      wins when its (normalised) overlap beats the best face overlap by more than 5 % (kEdgeHandicap), and the edge tests use
//      |R| + 1e-5 so that near-parallel edges never produce a spurious axis;
//   2. an edge axis gives ONE point: the closest point of B's edge to A's edge;
//   3. a face axis gives the incident face of the other box clipped to the reference face's rectangle (at most 8 vertices), the
//      vertices below the reference face kept, and when more than 4 remain the deepest one plus the three nearest to evenly
//      spaced directions around the polygon's centroid.
// Written for one lane per car pair: the axis loops are index arithmetic over (i, j), the clipping works on a small P2 polygon
// in registers / local memory.  out.normal points from box B towards box A, the points lie on B, depths are negative.
struct P2 { float u, v; RL_HDI float get(int k) const { return k ? v : u; } RL_HDI void set(int k, float x) { if (k) v = x; else u = x; } };

// One pass of polygon clipping against the half-plane  side * coord[k] < limit.  The output is cut off at 8 vertices (the caller
// then stops clipping altogether, like the reference does).  Returns the vertex count; `full` tells that the cut-off happened.
RL_HD inline int clip_against_halfplane(const P2* in, int nIn, P2* outPoly, int k, float side, float limit, bool& full) {
    int nOut = 0;
    full = false;
    for (int i = 0; i < nIn; i++) {
        const P2 cur = in[i], nxt = in[(i + 1 == nIn) ? 0 : i + 1];
        const bool curIn = side * cur.get(k) < limit, nxtIn = side * nxt.get(k) < limit;
        if (curIn) {
            outPoly[nOut++] = cur;
            if (nOut == 8) { full = true; return nOut; }
        }
        if (curIn != nxtIn) {  // the edge crosses the boundary: its intersection with coord[k] = side * limit
            P2 x;
            x.set(1 - k, cur.get(1 - k) + (nxt.get(1 - k) - cur.get(1 - k)) / (nxt.get(k) - cur.get(k)) * (side * limit - cur.get(k)));
            x.set(k, side * limit);
            outPoly[nOut++] = x;
            if (nOut == 8) { full = true; return nOut; }
        }
    }
    return nOut;
}

// quad (4 vertices) clipped to the rectangle |u| < hu, |v| < hv: -u, +u, -v, +v in turn
RL_HD inline int clip_quad_to_rect(float hu, float hv, const P2 quad[4], P2 result[8]) {
    P2 bufA[8], bufB[8];
    const P2* src = quad;
    int n = 4;
    P2* dst = bufA;
    for (int pass = 0; pass < 4; pass++) {
        const int k = pass >> 1;
        const float side = (pass & 1) ? 1.f : -1.f;
        bool full;
        n = clip_against_halfplane(src, n, dst, k, side, k ? hv : hu, full);
        src = dst;
        dst = (dst == bufA) ? bufB : bufA;
        if (full) break;
    }
    for (int i = 0; i < n; i++) result[i] = src[i];
    return n;
}

// Picks `want` of the n polygon vertices: `first` and then, for each of the remaining evenly spaced directions around the
// centroid (starting at `first`'s direction), the unused vertex whose direction is nearest.
RL_HD inline void pick_spread_vertices(int n, const P2* poly, int want, int first, int* chosen) {
    const float kPi_ = 3.14159265f;
    float cu, cv;
    if (n == 1) { cu = poly[0].u; cv = poly[0].v; }
    else if (n == 2) { cu = 0.5f * (poly[0].u + poly[1].u); cv = 0.5f * (poly[0].v + poly[1].v); }
    else {  // area-weighted centroid of the polygon (shoelace terms; the closing edge last)
        float area = 0, su = 0, sv = 0;
        for (int i = 0; i + 1 < n; i++) {
            const float w = poly[i].u * poly[i + 1].v - poly[i + 1].u * poly[i].v;
            area += w; su += w * (poly[i].u + poly[i + 1].u); sv += w * (poly[i].v + poly[i + 1].v);
        }
        const float w = poly[n - 1].u * poly[0].v - poly[0].u * poly[n - 1].v;
        const float scale = fabsf(area + w) > kEps ? 1.f / (3.0f * (area + w)) : 1e18f;
        cu = scale * (su + w * (poly[n - 1].u + poly[0].u));
        cv = scale * (sv + w * (poly[n - 1].v + poly[0].v));
    }
    float dir[8];
    unsigned freeMask = 0;
    for (int i = 0; i < n; i++) { dir[i] = rl_atan2(poly[i].v - cv, poly[i].u - cu); freeMask |= 1u << i; }
    freeMask &= ~(1u << first);
    chosen[0] = first;
    for (int j = 1; j < want; j++) {
        float target = (float)j * (2 * kPi_ / want) + dir[first];
        if (target > kPi_) target -= 2 * kPi_;
        int best = first;
        float bestGap = 1e9f;
        for (int i = 0; i < n; i++) {
            if (!((freeMask >> i) & 1u)) continue;
            float gap = fabsf(dir[i] - target);
            if (gap > kPi_) gap = 2 * kPi_ - gap;
            if (gap < bestGap) { bestGap = gap; best = i; }
        }
        freeMask &= ~(1u << best);
        chosen[j] = best;
    }
}

RL_HD RL_NOINLINE inline void box_box(V3 ca, const M3& ra, V3 ha, V3 cb, const M3& rb, V3 hb, BoxBoxResult& out) {
    out.n = 0;
    const float kEdgeHandicap = 1.05f, kParallelPad = 1.0e-5f;
    const V3 d = cb - ca;          // centre offset, world
    const V3 dA = tmul(d, ra);     // ... in A's frame
    float R[3][3], Q[3][3];        // R[i][j] = a_i . b_j, Q = |R|
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { R[i][j] = dot(ra.col(i), rb.col(j)); Q[i][j] = fabsf(R[i][j]); }

    // ---- 1. separating-axis search: code 1..3 = A's faces, 4..6 = B's faces, 7 + 3 i + j = edge a_i x b_j
    float best = -3.402823466e+38f;
    int code = 0;
    bool flip = false;
    V3 edgeAxisA(0, 0, 0);  // winning edge axis in A's frame (unit)
    for (int i = 0; i < 3; i++) {  // faces of A
        const float dist = dA[i], reach = ha[i] + hb[0] * Q[i][0] + hb[1] * Q[i][1] + hb[2] * Q[i][2];
        const float sep = fabsf(dist) - reach;
        if (sep > 0) return;
        if (sep > best) { best = sep; code = 1 + i; flip = dist < 0; }
    }
    for (int j = 0; j < 3; j++) {  // faces of B
        const float dist = dot(rb.col(j), d), reach = ha[0] * Q[0][j] + ha[1] * Q[1][j] + ha[2] * Q[2][j] + hb[j];
        const float sep = fabsf(dist) - reach;
        if (sep > 0) return;
        if (sep > best) { best = sep; code = 4 + j; flip = dist < 0; }
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q[i][j] += kParallelPad;
    for (int i = 0; i < 3; i++) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;          // a_i x b_j has no component along a_i
        const int il = i == 0 ? 1 : 0, ih = i == 2 ? 1 : 2;    // the other two axes of A, ascending
        for (int j = 0; j < 3; j++) {
            const int jl = j == 0 ? 1 : 0, jh = j == 2 ? 1 : 2;
            const float dist = dA[i2] * R[i1][j] - dA[i1] * R[i2][j];
            const float reach = ha[il] * Q[ih][j] + ha[ih] * Q[il][j] + hb[jl] * Q[i][jh] + hb[jh] * Q[i][jl];
            float sep = fabsf(dist) - reach;
            if (sep > kEps) return;
            V3 n(0, 0, 0);
            n[i1] = -R[i2][j]; n[i2] = R[i1][j];
            const float nl = sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
            if (nl > kEps) {
                sep /= nl;
                if (sep * kEdgeHandicap > best) { best = sep; code = 7 + 3 * i + j; flip = dist < 0; edgeAxisA = V3(n.x / nl, n.y / nl, n.z / nl); }
            }
        }
    }
    if (!code) return;
    V3 axis = code <= 3 ? ra.col(code - 1) : (code <= 6 ? rb.col(code - 4) : ra * edgeAxisA);  // from A towards B
    if (flip) axis = -axis;
    const float overlap = -best;
    out.normal = -axis;

    // ---- 2. edge-edge: one point, on B's edge
    if (code > 6) {
        V3 onA = ca, onB = cb;  // the supporting vertices along the axis
        for (int k = 0; k < 3; k++) onA += ra.col(k) * ((dot(axis, ra.col(k)) > 0 ? 1.f : -1.f) * ha[k]);
        for (int k = 0; k < 3; k++) onB += rb.col(k) * ((dot(axis, rb.col(k)) > 0 ? -1.f : 1.f) * hb[k]);
        const V3 ea = ra.col((code - 7) / 3), eb = rb.col((code - 7) % 3);
        const V3 gap = onB - onA;
        const float cosAB = dot(ea, eb), alongA = dot(ea, gap), alongB = -dot(eb, gap);
        float denom = 1 - cosAB * cosAB, tB = 0;
        if (denom > 0.0001f) { denom = 1.f / denom; tB = (cosAB * alongA + alongB) * denom; }
        out.point[0] = onB + eb * tB; out.depth[0] = -overlap; out.n = 1;
        return;
    }

    // ---- 3. face contact: the reference face belongs to the box that owns the axis, the incident face to the other box
    const bool refIsA = code <= 3;
    const M3 &Rr = refIsA ? ra : rb, &Ri = refIsA ? rb : ra;
    const V3 cr = refIsA ? ca : cb, ci = refIsA ? cb : ca;
    const V3 hr = refIsA ? ha : hb, hi = refIsA ? hb : ha;
    const V3 outward = refIsA ? axis : -axis;      // reference face normal, pointing at the incident box
    const V3 inInc = tmul(outward, Ri);            // ... in the incident box's frame
    const V3 mag = vabs(inInc);
    int face, s1, s2;  // incident face axis = the most anti-parallel one; s1 < s2 span the face
    if (mag[1] > mag[0]) {
        if (mag[1] > mag[2]) { face = 1; s1 = 0; s2 = 2; } else { face = 2; s1 = 0; s2 = 1; }
    } else {
        if (mag[0] > mag[2]) { face = 0; s1 = 1; s2 = 2; } else { face = 2; s1 = 0; s2 = 1; }
    }
    V3 centre = ci - cr;  // incident face centre relative to the reference box
    if (inInc[face] < 0) centre = centre + Ri.col(face) * hi[face]; else centre = centre - Ri.col(face) * hi[face];
    const int refAxis = refIsA ? code - 1 : code - 4;
    const int t1 = refAxis == 0 ? 1 : 0, t2 = refAxis == 2 ? 1 : 2;  // the reference face's in-plane axes
    // incident face corners in the reference face's 2-D frame
    const float cu = dot(centre, Rr.col(t1)), cv = dot(centre, Rr.col(t2));
    float j11 = dot(Rr.col(t1), Ri.col(s1)), j12 = dot(Rr.col(t1), Ri.col(s2));
    float j21 = dot(Rr.col(t2), Ri.col(s1)), j22 = dot(Rr.col(t2), Ri.col(s2));
    P2 quad[4];
    {
        const float e1u = j11 * hi[s1], e1v = j21 * hi[s1], e2u = j12 * hi[s2], e2v = j22 * hi[s2];
        quad[0].u = cu - e1u - e2u; quad[0].v = cv - e1v - e2v;
        quad[1].u = cu - e1u + e2u; quad[1].v = cv - e1v + e2v;
        quad[2].u = cu + e1u + e2u; quad[2].v = cv + e1v + e2v;
        quad[3].u = cu + e1u - e2u; quad[3].v = cv + e1v - e2v;
    }
    P2 poly[8];
    const int nClip = clip_quad_to_rect(hr[t1], hr[t2], quad, poly);
    if (nClip < 1) return;
    // back to 3-D through the inverse of the 2x2 projection, keeping the vertices that are below the reference face
    const float invDet = 1.f / (j11 * j22 - j12 * j21);
    j11 *= invDet; j12 *= invDet; j21 *= invDet; j22 *= invDet;
    V3 pts[8];
    float deep[8];
    int kept = 0;
    for (int v = 0; v < nClip; v++) {
        const float a1 = j22 * (poly[v].u - cu) - j12 * (poly[v].v - cv);
        const float a2 = -j21 * (poly[v].u - cu) + j11 * (poly[v].v - cv);
        pts[kept] = centre + Ri.col(s1) * a1 + Ri.col(s2) * a2;
        deep[kept] = hr[refAxis] - dot(outward, pts[kept]);
        if (deep[kept] >= 0) { poly[kept] = poly[v]; kept++; }
    }
    if (kept < 1) return;
    int order[8];
    int nOut = kept;
    if (kept <= 4) {
        for (int v = 0; v < kept; v++) order[v] = v;
    } else {
        int deepest = 0;
        for (int v = 1; v < kept; v++) if (deep[v] > deep[deepest]) deepest = v;
        pick_spread_vertices(kept, poly, 4, deepest, order);
        nOut = 4;
    }
    for (int v = 0; v < nOut; v++) {
        const int q = order[v];
        V3 w = pts[q] + cr;
        if (!refIsA) w = w - axis * deep[q];  // the points are reported on B: project off A's ... off the reference face when B owns it
        out.point[out.n] = w; out.depth[out.n] = -deep[q]; out.n++;
    }
}

}  // namespace rl
