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

// ---- btBoxBoxDetector (ODE dBoxBox2; B/BulletCollision/CollisionDispatch/btBoxBoxDetector.cpp:260-720) ----------
// 15-axis separating-axis search (edge axes handicapped by fudge 1.05), then either one edge-edge point or the
// incident face clipped against the reference face, culled to at most 4 points.  Contacts only when the full
// (with-margin) boxes interpenetrate.  out.normal is normalWorldOnB (from box B towards box A), points lie on B.
RL_HD inline int bb_clip_rect_quad(const float h[2], const float p[8], float ret[16]) {
    int nq = 4, nr = 0;
    float buffer[16];
    const float* q = p;
    float* r = ret;
    for (int dir = 0; dir <= 1; dir++) {
        for (int sign = -1; sign <= 1; sign += 2) {
            const float* pq = q;
            float* pr = r;
            nr = 0;
            for (int i = nq; i > 0; i--) {
                if (sign * pq[dir] < h[dir]) {
                    pr[0] = pq[0]; pr[1] = pq[1]; pr += 2; nr++;
                    if (nr & 8) { q = r; goto done; }
                }
                const float* nextq = (i > 1) ? pq + 2 : q;
                if ((sign * pq[dir] < h[dir]) ^ (sign * nextq[dir] < h[dir])) {
                    pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sign * h[dir] - pq[dir]);
                    pr[dir] = sign * h[dir];
                    pr += 2; nr++;
                    if (nr & 8) { q = r; goto done; }
                }
                pq += 2;
            }
            q = r;
            r = (q == ret) ? buffer : ret;
            nq = nr;
        }
    }
done:
    if (q != ret) for (int i = 0; i < nr * 2; i++) ret[i] = q[i];
    return nr;
}

RL_HD inline void bb_cull_points(int n, const float* p, int m, int i0, int* iret) {
    const float PI_ = 3.14159265f;
    float a, cx, cy, q;
    if (n == 1) { cx = p[0]; cy = p[1]; }
    else if (n == 2) { cx = 0.5f * (p[0] + p[2]); cy = 0.5f * (p[1] + p[3]); }
    else {
        a = 0; cx = 0; cy = 0;
        for (int i = 0; i < (n - 1); i++) {
            q = p[i * 2] * p[i * 2 + 3] - p[i * 2 + 2] * p[i * 2 + 1];
            a += q; cx += q * (p[i * 2] + p[i * 2 + 2]); cy += q * (p[i * 2 + 1] + p[i * 2 + 3]);
        }
        q = p[n * 2 - 2] * p[1] - p[0] * p[n * 2 - 1];
        if (fabsf(a + q) > kEps) a = 1.f / (3.0f * (a + q)); else a = 1e18f;
        cx = a * (cx + q * (p[n * 2 - 2] + p[0]));
        cy = a * (cy + q * (p[n * 2 - 1] + p[1]));
    }
    float A[8];
    int avail[8];
    for (int i = 0; i < n; i++) { A[i] = rl_atan2(p[i * 2 + 1] - cy, p[i * 2] - cx); avail[i] = 1; }
    avail[i0] = 0;
    iret[0] = i0;
    iret++;
    for (int j = 1; j < m; j++) {
        a = (float)j * (2 * PI_ / m) + A[i0];
        if (a > PI_) a -= 2 * PI_;
        float maxdiff = 1e9f, diff;
        *iret = i0;
        for (int i = 0; i < n; i++) {
            if (avail[i]) {
                diff = fabsf(A[i] - a);
                if (diff > PI_) diff = 2 * PI_ - diff;
                if (diff < maxdiff) { maxdiff = diff; *iret = i; }
            }
        }
        avail[*iret] = 0;
        iret++;
    }
}

RL_HD RL_NOINLINE inline void box_box(V3 p1, const M3& R1, V3 A, V3 p2, const M3& R2, V3 B, BoxBoxResult& out) {
    out.n = 0;
    const float fudge_factor = 1.05f;
    V3 p = p2 - p1;
    V3 pp = tmul(p, R1);
    float R[3][3], Q[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { R[i][j] = dot(R1.col(i), R2.col(j)); Q[i][j] = fabsf(R[i][j]); }
    float s = -3.402823466e+38f, s2, l;
    int invert_normal = 0, code = 0;
    int normalRBox = 0, normalRCol = 0;  // which box / column the best face axis comes from
    bool normalIsFace = false;
    V3 normalC(0, 0, 0);
#define RL_TST_FACE(expr1, expr2, box, col, cc) \
    s2 = fabsf(expr1) - (expr2);                \
    if (s2 > 0) return;                         \
    if (s2 > s) { s = s2; normalIsFace = true; normalRBox = box; normalRCol = col; invert_normal = ((expr1) < 0); code = (cc); }
    RL_TST_FACE(pp[0], (A[0] + B[0] * Q[0][0] + B[1] * Q[0][1] + B[2] * Q[0][2]), 1, 0, 1);
    RL_TST_FACE(pp[1], (A[1] + B[0] * Q[1][0] + B[1] * Q[1][1] + B[2] * Q[1][2]), 1, 1, 2);
    RL_TST_FACE(pp[2], (A[2] + B[0] * Q[2][0] + B[1] * Q[2][1] + B[2] * Q[2][2]), 1, 2, 3);
    RL_TST_FACE(dot(R2.col(0), p), (A[0] * Q[0][0] + A[1] * Q[1][0] + A[2] * Q[2][0] + B[0]), 2, 0, 4);
    RL_TST_FACE(dot(R2.col(1), p), (A[0] * Q[0][1] + A[1] * Q[1][1] + A[2] * Q[2][1] + B[1]), 2, 1, 5);
    RL_TST_FACE(dot(R2.col(2), p), (A[0] * Q[0][2] + A[1] * Q[1][2] + A[2] * Q[2][2] + B[2]), 2, 2, 6);
#undef RL_TST_FACE
#define RL_TST_EDGE(expr1, expr2, n1, n2, n3, cc)          \
    s2 = fabsf(expr1) - (expr2);                           \
    if (s2 > kEps) return;                                 \
    l = sqrtf((n1) * (n1) + (n2) * (n2) + (n3) * (n3));    \
    if (l > kEps) {                                        \
        s2 /= l;                                           \
        if (s2 * fudge_factor > s) { s = s2; normalIsFace = false; normalC = V3((n1) / l, (n2) / l, (n3) / l); invert_normal = ((expr1) < 0); code = (cc); } \
    }
    const float fudge2 = 1.0e-5f;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q[i][j] += fudge2;
    RL_TST_EDGE(pp[2] * R[1][0] - pp[1] * R[2][0], (A[1] * Q[2][0] + A[2] * Q[1][0] + B[1] * Q[0][2] + B[2] * Q[0][1]), 0.f, -R[2][0], R[1][0], 7);
    RL_TST_EDGE(pp[2] * R[1][1] - pp[1] * R[2][1], (A[1] * Q[2][1] + A[2] * Q[1][1] + B[0] * Q[0][2] + B[2] * Q[0][0]), 0.f, -R[2][1], R[1][1], 8);
    RL_TST_EDGE(pp[2] * R[1][2] - pp[1] * R[2][2], (A[1] * Q[2][2] + A[2] * Q[1][2] + B[0] * Q[0][1] + B[1] * Q[0][0]), 0.f, -R[2][2], R[1][2], 9);
    RL_TST_EDGE(pp[0] * R[2][0] - pp[2] * R[0][0], (A[0] * Q[2][0] + A[2] * Q[0][0] + B[1] * Q[1][2] + B[2] * Q[1][1]), R[2][0], 0.f, -R[0][0], 10);
    RL_TST_EDGE(pp[0] * R[2][1] - pp[2] * R[0][1], (A[0] * Q[2][1] + A[2] * Q[0][1] + B[0] * Q[1][2] + B[2] * Q[1][0]), R[2][1], 0.f, -R[0][1], 11);
    RL_TST_EDGE(pp[0] * R[2][2] - pp[2] * R[0][2], (A[0] * Q[2][2] + A[2] * Q[0][2] + B[0] * Q[1][1] + B[1] * Q[1][0]), R[2][2], 0.f, -R[0][2], 12);
    RL_TST_EDGE(pp[1] * R[0][0] - pp[0] * R[1][0], (A[0] * Q[1][0] + A[1] * Q[0][0] + B[1] * Q[2][2] + B[2] * Q[2][1]), -R[1][0], R[0][0], 0.f, 13);
    RL_TST_EDGE(pp[1] * R[0][1] - pp[0] * R[1][1], (A[0] * Q[1][1] + A[1] * Q[0][1] + B[0] * Q[2][2] + B[2] * Q[2][0]), -R[1][1], R[0][1], 0.f, 14);
    RL_TST_EDGE(pp[1] * R[0][2] - pp[0] * R[1][2], (A[0] * Q[1][2] + A[1] * Q[0][2] + B[0] * Q[2][1] + B[1] * Q[2][0]), -R[1][2], R[0][2], 0.f, 15);
#undef RL_TST_EDGE
    if (!code) return;
    V3 normal;
    if (normalIsFace) normal = (normalRBox == 1 ? R1 : R2).col(normalRCol);
    else normal = R1 * normalC;
    if (invert_normal) normal = -normal;
    float depth = -s;
    out.normal = -normal;
    if (code > 6) {
        V3 pa = p1;
        for (int j = 0; j < 3; j++) { float sign = (dot(normal, R1.col(j)) > 0) ? 1.f : -1.f; pa += R1.col(j) * (sign * A[j]); }
        V3 pb = p2;
        for (int j = 0; j < 3; j++) { float sign = (dot(normal, R2.col(j)) > 0) ? -1.f : 1.f; pb += R2.col(j) * (sign * B[j]); }
        V3 ua = R1.col((code - 7) / 3), ub = R2.col((code - 7) % 3);
        // dLineClosestApproach
        V3 dpp = pb - pa;
        float uaub = dot(ua, ub), q1 = dot(ua, dpp), q2 = -dot(ub, dpp);
        float d = 1 - uaub * uaub, alpha, beta;
        if (d <= 0.0001f) { alpha = 0; beta = 0; }
        else { d = 1.f / d; alpha = (q1 + uaub * q2) * d; beta = (uaub * q1 + q2) * d; }
        (void)alpha;
        pb += ub * beta;
        out.point[0] = pb; out.depth[0] = -depth; out.n = 1;
        return;
    }
    const M3 &Ra = code <= 3 ? R1 : R2, &Rb = code <= 3 ? R2 : R1;
    V3 pa = code <= 3 ? p1 : p2, pb = code <= 3 ? p2 : p1;
    V3 Sa = code <= 3 ? A : B, Sb = code <= 3 ? B : A;
    V3 normal2 = code <= 3 ? normal : -normal;
    V3 nr = tmul(normal2, Rb);
    V3 anr = vabs(nr);
    int lanr, a1, a2;
    if (anr[1] > anr[0]) {
        if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
    } else {
        if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
    }
    V3 center;
    if (nr[lanr] < 0) center = pb - pa + Rb.col(lanr) * Sb[lanr];
    else center = pb - pa - Rb.col(lanr) * Sb[lanr];
    int codeN = code <= 3 ? code - 1 : code - 4, code1, code2;
    if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
    float quad[8];
    float c1 = dot(center, Ra.col(code1)), c2 = dot(center, Ra.col(code2));
    float m11 = dot(Ra.col(code1), Rb.col(a1)), m12 = dot(Ra.col(code1), Rb.col(a2));
    float m21 = dot(Ra.col(code2), Rb.col(a1)), m22 = dot(Ra.col(code2), Rb.col(a2));
    {
        float k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
        quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
        quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
        quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
        quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
    }
    float rect[2] = {Sa[code1], Sa[code2]};
    float ret[16];
    int n = bb_clip_rect_quad(rect, quad, ret);
    if (n < 1) return;
    V3 point[8];
    float dep[8];
    float det1 = 1.f / (m11 * m22 - m12 * m21);
    m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
    int cnum = 0;
    for (int j = 0; j < n; j++) {
        float k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
        float k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
        point[cnum] = center + Rb.col(a1) * k1 + Rb.col(a2) * k2;
        dep[cnum] = Sa[codeN] - dot(normal2, point[cnum]);
        if (dep[cnum] >= 0) { ret[cnum * 2] = ret[j * 2]; ret[cnum * 2 + 1] = ret[j * 2 + 1]; cnum++; }
    }
    if (cnum < 1) return;
    int maxc = 4;
    if (maxc > cnum) maxc = cnum;
    if (maxc < 1) maxc = 1;
    if (cnum <= maxc) {
        for (int j = 0; j < cnum; j++) {
            V3 w = point[j] + pa;
            if (code >= 4) w = w - normal * dep[j];
            out.point[out.n] = w; out.depth[out.n] = -dep[j]; out.n++;
        }
    } else {
        int i1 = 0;
        float maxdepth = dep[0];
        for (int i = 1; i < cnum; i++) if (dep[i] > maxdepth) { maxdepth = dep[i]; i1 = i; }
        int iret[8];
        bb_cull_points(cnum, ret, maxc, i1, iret);
        for (int j = 0; j < maxc; j++) {
            V3 w = point[iret[j]] + pa;
            if (code >= 4) w = w - normal * dep[iret[j]];
            out.point[out.n] = w; out.depth[out.n] = -dep[iret[j]]; out.n++;
        }
    }
}

}  // namespace rl
