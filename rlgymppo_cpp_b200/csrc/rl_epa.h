// rl_epa.h — penetration depth of two overlapping convex shapes: the expanding-polytope search the reference falls back to
// when btGjkPairDetector finds the margin-less cores touching or overlapping (btGjkPairDetector.cpp:860-940 ->
// btGjkEpaPenetrationDepthSolver::calcPenDepth -> btGjkEpaSolver2::Penetration / ::Distance, btGjkEpa2.cpp).
//
// What has to be reproduced for contact parity, and is:
//  * the search runs on the shapes WITH their margins, in shape A's local frame: the hitbox is the margin-rounded box
//    (core corner + margin * unit direction, btConvexShape::localGetSupportVertexNonVirtual), the triangle keeps margin 0;
//  * its own GJK (projectorigin on 2/3/4 points, the 4-entry duplicate ring, the omega/alpha exit), the origin-enclosing
//    completion of a degenerate simplex, the tetrahedron orientation, face distances measured to the closest FEATURE of a
//    face (plane, edge or vertex: getedgedist), best-face selection in hull-list order, the horizon expansion order, the
//    1e-4 accuracy exit, the 128-vertex / 256-face limits and the fall-back results;
//  * the nine guess directions of calcPenDepth and the margin-less Distance query after a failed Penetration;
//  * the quirk that sResults::normal is expressed in A's LOCAL frame (btGjkEpa2.cpp:1006, :955).
//
// Own structure: everything is index based (no pointers), vertices of the GJK simplex and of the polytope share one
// store, faces live in two index-linked lists (hull / stock) that hand faces out in the reference's order, and the
// horizon walk is an explicit stack instead of recursion (device code: no unbounded call depth).  One evaluation needs
// an EpaWs workspace (~13 KB): the host build puts it on the stack, the role kernel keeps one per warp in global memory
// (contacts this deep are rare; lanes take turns — engine.cu).
#pragma once
#include <stddef.h>
#include "rl_math.h"

namespace rl {

#if defined(RL_DEBUG_CONTACTS) && !defined(__CUDA_ARCH__)
static long g_dbg_epa_overflows = 0;
static long g_dbg_epa_calls = 0, g_dbg_epa_iters = 0, g_dbg_epa_maxface = 0, g_dbg_epa_maxsv = 0, g_dbg_epa_hist[8] = {0};  // host debugging only
#endif
#if defined(RLG_EPA_TIMING) && defined(__CUDACC__)
// diagnostic builds: [0] cycles inside the search, [1] calls, [2] cycles in GJK::Evaluate, [3] GJK evaluations, [4] GJK iterations,
// [5] cycles in EPA::Evaluate, [6] EPA iterations, [7] guesses tried
static __device__ unsigned long long g_epa_timing[8];
#define EPA_T0() const long long epaT0_ = clock64()
#define EPA_T(i) atomicAdd(&g_epa_timing[i], (unsigned long long)(clock64() - epaT0_))
#define EPA_N(i, n) atomicAdd(&g_epa_timing[i], (unsigned long long)(n))
#else
#define EPA_T0() do {} while (0)
#define EPA_T(i) do {} while (0)
#define EPA_N(i, n) do {} while (0)
#endif
constexpr int kEpaMaxVerts = 128, kEpaMaxFaces = 256, kEpaMaxIter = 255, kGjk2MaxIter = 128;
constexpr float kGjk2Accuracy = 1e-4f, kGjk2MinDist = 1e-4f, kGjk2DupEps = 1e-4f;
constexpr float kEpaAccuracy = 1e-4f, kEpaPlaneEps = 1e-5f;

struct EpaSV { V3 d, w; RL_HDI EpaSV() : d(NoInit()), w(NoInit()) {} };
struct EpaFace {
    V3 n; float d;
    int16_t f[3];  // neighbour across edge i
    int16_t l[2];  // list links: previous, next
    uint8_t c[3];  // vertices (store indices)
    uint8_t e[3];  // the neighbour's edge that is bound to edge i
    uint8_t pass;
    RL_HDI EpaFace() : n(NoInit()) {}
};
// A workspace is a view of raw storage with capacities.  Full capacity (128 vertices / 256 faces) is the reference's; the
// role kernel first runs in a SMALL workspace in shared memory (every evaluation seen in play fits: the high-water marks over
// 600 k car-ticks of random play and the scripted wall / goal / ramp hits are 19 faces and 8 vertices) and repeats the
// evaluation in the full-size one only if the small one runs out — the algorithm is deterministic, so a run that fits is the
// same run.
struct EpaWs {
    EpaSV* sv;        // [4 + maxVerts]: [0,4) the GJK simplex store, then the polytope vertices
    EpaFace* fc;      // [maxFaces]
    uint16_t* stack;  // [2 * maxFaces] horizon walk frames: (face << 4) | (edge << 2) | stage
    int32_t maxVerts, maxFaces;
};
constexpr int kEpaSmallVerts = 10, kEpaSmallFaces = 24;
RL_HDI constexpr size_t epa_ws_bytes(int verts, int faces) { return sizeof(EpaSV) * (4 + verts) + sizeof(EpaFace) * faces + 2 * 2 * faces; }
// a workspace block = 16 bytes (lock word + padding) followed by the storage
constexpr size_t kEpaSmallBytes = 16 + (epa_ws_bytes(kEpaSmallVerts, kEpaSmallFaces) + 15) / 16 * 16;
constexpr size_t kEpaFullBytes = 16 + (epa_ws_bytes(kEpaMaxVerts, kEpaMaxFaces) + 15) / 16 * 16;

RL_HDI EpaWs epa_ws_view(void* mem, int verts, int faces) {  // mem: 4-byte aligned, epa_ws_bytes(verts, faces) long
    EpaWs w;
    w.sv = reinterpret_cast<EpaSV*>(mem);
    w.fc = reinterpret_cast<EpaFace*>(w.sv + 4 + verts);
    w.stack = reinterpret_cast<uint16_t*>(w.fc + faces);
    w.maxVerts = verts; w.maxFaces = faces;
    return w;
}
#if !defined(__CUDA_ARCH__)
static int g_epa_host_small_verts = kEpaSmallVerts, g_epa_host_small_faces = kEpaSmallFaces;  // tests shrink them to force the fallback
#endif
// What a caller hands down (device): the block's small workspace in shared memory and its full-size one in global memory,
// each behind a lock word (lanes take turns; deep contacts are rare: ~5e-4 per car-tick in random play).  Host: nullptr = a local full-size workspace.
struct EpaCtx {
    int32_t* smallLock; void* smallMem;
    int32_t* fullLock; void* fullMem;
};
RL_HDI EpaCtx epa_ctx(void* smallBlock, void* fullBlock) {
    EpaCtx c;
    c.smallLock = reinterpret_cast<int32_t*>(smallBlock); c.smallMem = reinterpret_cast<unsigned char*>(smallBlock) + 16;
    c.fullLock = reinterpret_cast<int32_t*>(fullBlock); c.fullMem = reinterpret_cast<unsigned char*>(fullBlock) + 16;
    return c;
}

// The Minkowski difference A - B in A's local frame (gjkepa2_impl::MinkowskiDiff).  A = box core (+ margin sphere),
// B = any convex given by supB(dirInB, withMargin) in B's local frame.
template <class SupB>
struct Mink {
    M3 rotA;          // wtrs0 basis; B's basis is the identity (static world triangles, the ball's point core), so
                      // toshape1 = basisB^T * basisA = rotA and toshape0's basis = rotA^T
    V3 originA;       // wtrs0 origin (already shifted by the pair detector's position offset)
    V3 rel;           // toshape0 origin: (originB - originA) * rotA
    V3 coreHalf; float marginA;
    SupB supB;        // supB(dirInB, withMargin) in B's local frame
    bool margins;
    RL_HDI Mink(const M3& r, V3 oA, V3 oB, V3 core, float mA, SupB sb) : rotA(r), originA(oA), rel(tmul(oB - oA, r)), coreHalf(core), marginA(mA), supB(sb), margins(true) {}
    RL_HDI V3 support0(V3 d) const {
        if (margins) {
            V3 dn = d;
            if (len2(dn) < kEps * kEps) dn = V3(-1, -1, -1);
            dn = normalized(dn);
            return V3(dn.x >= 0 ? coreHalf.x : -coreHalf.x, dn.y >= 0 ? coreHalf.y : -coreHalf.y, dn.z >= 0 ? coreHalf.z : -coreHalf.z) + dn * marginA;
        }
        return V3(d.x >= 0 ? coreHalf.x : -coreHalf.x, d.y >= 0 ? coreHalf.y : -coreHalf.y, d.z >= 0 ? coreHalf.z : -coreHalf.z);
    }
    RL_HDI V3 support1(V3 d) const { return tmul(supB(rotA * d, margins), rotA) + rel; }
    RL_HDI V3 support(V3 d) const { return support0(d) - support1(-d); }
    RL_HDI V3 to_world(V3 p) const { return rotA * p + originA; }
};

// (Measured, profiles/r02o_epa_outofline.txt: the search costs ~104 k cycles per call, 50-90 % of its stall samples are instruction
// fetch — one lane running cold, branchy code pays every fetch alone.  Making every helper with several call sites a real function
// shrank k_roles from 907 to 737 KB but made a call SLOWER (117 k cycles, k_roles 1.26 -> 1.32 ms): the calls spill the live state
// of a 168-register kernel.  So the helpers stay inlined.)
// One evaluation at a time per workspace.  f(ws) returns false when the workspace ran out of room (small workspace only).
// Every entry of a workspace is written before it is read within one evaluation, so nothing of the previous holder is consumed.
#if defined(__CUDA_ARCH__)
template <class F>
RL_HDI bool epa_locked(int32_t* lock, F f) {
    bool done = false, ok = false;
    while (!done) {
        if (atomicCAS(lock, 0, 1) == 0) {
            ok = f();
            __threadfence_block();
            atomicExch(lock, 0);
            done = true;
        }
    }
    return ok;
}
#endif
template <class F>
RL_HDI void with_epa_ws(const EpaCtx* ctx, F f) {
#if defined(RLG_NO_EPA)  // A/B builds only: what the penetration-depth search costs
    (void)ctx; (void)f;
#elif defined(__CUDA_ARCH__)
    // 1. the block's shared-memory workspace when nobody holds it — no waiting: a car that digs into the mesh overlaps several
    //    triangles at once, its pairs sit on neighbouring lanes of one warp, and lanes queueing for one workspace would run their
    //    searches one after the other with the waiting lanes spinning next to the working one (measured: 100 k cycles per call);
    // 2. otherwise a private small workspace on the lane's own stack: those lanes search side by side in the same code;
    // 3. the full-size workspace of the block (global memory, behind its lock) in the rare case the small capacity runs out.
    bool ok = false;
    if (atomicCAS(ctx->smallLock, 0, 1) == 0) {
        EpaWs w = epa_ws_view(ctx->smallMem, kEpaSmallVerts, kEpaSmallFaces);
        ok = f(w);
        __threadfence_block();
        atomicExch(ctx->smallLock, 0);
    } else {
        alignas(16) unsigned char mem[(epa_ws_bytes(kEpaSmallVerts, kEpaSmallFaces) + 15) / 16 * 16];
        EpaWs w = epa_ws_view(mem, kEpaSmallVerts, kEpaSmallFaces);
        ok = f(w);
    }
    if (ok) return;
    epa_locked(ctx->fullLock, [&]() { EpaWs w = epa_ws_view(ctx->fullMem, kEpaMaxVerts, kEpaMaxFaces); return f(w); });
#else
    (void)ctx;  // host test build: the same two-step protocol on local storage
    static thread_local unsigned char mem[kEpaFullBytes];
    EpaWs w = epa_ws_view(mem, g_epa_host_small_verts, g_epa_host_small_faces);
    if (f(w)) return;
#if defined(RL_DEBUG_CONTACTS)
    g_dbg_epa_overflows++;
#endif
    w = epa_ws_view(mem, kEpaMaxVerts, kEpaMaxFaces);
    f(w);
#endif
}

struct Gjk2Simplex { uint8_t c[4]; float p[4]; int rank; };
struct Gjk2 {
    Gjk2Simplex sx[2];
    uint8_t freeList[4];
    int nfree, current, status;  // status: 0 valid, 1 inside, 2 failed
    V3 ray;
    float distance;
};

RL_HDI float det3(V3 a, V3 b, V3 c) {
    return a.y * b.z * c.x + a.z * b.x * c.y - a.x * b.z * c.y - a.y * b.x * c.z + a.x * b.y * c.z - a.z * b.y * c.x;
}

// GJK::projectorigin, 2 / 3 / 4 points: squared distance of the origin to the simplex, barycentric weights, used-vertex mask
RL_HD inline float project_origin2(V3 a, V3 b, float* w, unsigned& m) {
    const V3 d = b - a;
    const float l = len2(d);
    if (l > 0.f) {
        const float t = l > 0 ? -dot(a, d) / l : 0;
        if (t >= 1) { w[0] = 0; w[1] = 1; m = 2; return len2(b); }
        else if (t <= 0) { w[0] = 1; w[1] = 0; m = 1; return len2(a); }
        else { w[0] = 1 - (w[1] = t); m = 3; return len2(a + d * t); }
    }
    return -1;
}
RL_HD inline float project_origin3(V3 a, V3 b, V3 c, float* w, unsigned& m) {
    const V3 vt[3] = {a, b, c};
    const V3 dl[3] = {a - b, b - c, c - a};
    const V3 n = cross(dl[0], dl[1]);
    const float l = len2(n);
    if (l > 0.f) {
        float mindist = -1;
        float subw[2] = {0.f, 0.f};
        unsigned subm = 0;
        for (int i = 0; i < 3; ++i) {
            if (dot(vt[i], cross(dl[i], n)) > 0) {
                const int j = (i + 1) % 3;
                const float subd = project_origin2(vt[i], vt[j], subw, subm);
                if (mindist < 0 || subd < mindist) {
                    mindist = subd;
                    m = ((subm & 1) ? 1u << i : 0u) + ((subm & 2) ? 1u << j : 0u);
                    w[i] = subw[0]; w[j] = subw[1]; w[(j + 1) % 3] = 0;
                }
            }
        }
        if (mindist < 0) {
            const float d = dot(a, n);
            const float s = sqrtf(l);
            const V3 p = n * (d / l);
            mindist = len2(p);
            m = 7;
            w[0] = len(cross(dl[1], b - p)) / s;
            w[1] = len(cross(dl[2], c - p)) / s;
            w[2] = 1 - (w[0] + w[1]);
        }
        return mindist;
    }
    return -1;
}
RL_HD inline float project_origin4(V3 a, V3 b, V3 c, V3 d, float* w, unsigned& m) {
    const V3 vt[4] = {a, b, c, d};
    const V3 dl[3] = {a - d, b - d, c - d};
    const float vl = det3(dl[0], dl[1], dl[2]);
    const bool ng = (vl * dot(a, cross(b - c, a - b))) <= 0;
    if (ng && fabsf(vl) > 0.f) {
        float mindist = -1;
        float subw[3] = {0.f, 0.f, 0.f};
        unsigned subm = 0;
        for (int i = 0; i < 3; ++i) {
            const int j = (i + 1) % 3;
            const float s = vl * dot(d, cross(dl[i], dl[j]));
            if (s > 0) {
                const float subd = project_origin3(vt[i], vt[j], d, subw, subm);
                if (mindist < 0 || subd < mindist) {
                    mindist = subd;
                    m = ((subm & 1) ? 1u << i : 0u) + ((subm & 2) ? 1u << j : 0u) + ((subm & 4) ? 8u : 0u);
                    w[i] = subw[0]; w[j] = subw[1]; w[(j + 1) % 3] = 0; w[3] = subw[2];
                }
            }
        }
        if (mindist < 0) {
            mindist = 0;
            m = 15;
            w[0] = det3(c, b, d) / vl;
            w[1] = det3(a, c, d) / vl;
            w[2] = det3(b, a, d) / vl;
            w[3] = 1 - (w[0] + w[1] + w[2]);
        }
        return mindist;
    }
    return -1;
}

template <class Sh>
RL_HDI void gjk2_support(const Sh& sh, V3 d, EpaSV& sv) {
    sv.d = d / len(d);
    sv.w = sh.support(sv.d);
}
template <class Sh>
RL_HDI void gjk2_append(Gjk2& g, EpaSV* store, const Sh& sh, Gjk2Simplex& s, V3 v) {
    s.p[s.rank] = 0;
    s.c[s.rank] = g.freeList[--g.nfree];
    gjk2_support(sh, v, store[s.c[s.rank++]]);
}
RL_HDI void gjk2_remove(Gjk2& g, Gjk2Simplex& s) { g.freeList[g.nfree++] = s.c[--s.rank]; }

// GJK::Evaluate (btGjkEpa2.cpp:196-333)
template <class Sh>
RL_HD inline int gjk2_evaluate(Gjk2& g, EpaSV* store, const Sh& sh, V3 guess) {
    int iterations = 0;
    float sqdist = 0, alpha = 0;
#if defined(__CUDA_ARCH__)
    EPA_T0();
#endif
    V3 lastw[4];
    int clastw = 0;
    for (int i = 0; i < 4; i++) g.freeList[i] = (uint8_t)i;
    g.nfree = 4; g.current = 0; g.status = 0; g.distance = 0;
    g.sx[0].rank = 0;
    g.ray = guess;
    const float sqrl = len2(g.ray);
    gjk2_append(g, store, sh, g.sx[0], sqrl > 0 ? -g.ray : V3(1, 0, 0));
    g.sx[0].p[0] = 1;
    g.ray = store[g.sx[0].c[0]].w;
    sqdist = sqrl;
    lastw[0] = lastw[1] = lastw[2] = lastw[3] = g.ray;
    do {
        const int next = 1 - g.current;
        Gjk2Simplex& cs = g.sx[g.current];
        Gjk2Simplex& ns = g.sx[next];
        const float rl = len(g.ray);
        if (rl < kGjk2MinDist) { g.status = 1; break; }
        gjk2_append(g, store, sh, cs, -g.ray);
        const V3 w = store[cs.c[cs.rank - 1]].w;
        bool found = false;
        for (int i = 0; i < 4; ++i)
            if (len2(w - lastw[i]) < kGjk2DupEps) { found = true; break; }
        if (found) { gjk2_remove(g, cs); break; }
        lastw[clastw = (clastw + 1) & 3] = w;
        const float omega = dot(g.ray, w) / rl;
        alpha = fmaxf_(omega, alpha);
        if (((rl - alpha) - (kGjk2Accuracy * rl)) <= 0) { gjk2_remove(g, cs); break; }
        float weights[4];
        unsigned mask = 0;
        switch (cs.rank) {
        case 2: sqdist = project_origin2(store[cs.c[0]].w, store[cs.c[1]].w, weights, mask); break;
        case 3: sqdist = project_origin3(store[cs.c[0]].w, store[cs.c[1]].w, store[cs.c[2]].w, weights, mask); break;
        case 4: sqdist = project_origin4(store[cs.c[0]].w, store[cs.c[1]].w, store[cs.c[2]].w, store[cs.c[3]].w, weights, mask); break;
        }
        if (sqdist >= 0) {
            ns.rank = 0;
            g.ray = V3(0, 0, 0);
            g.current = next;
            for (int i = 0, ni = cs.rank; i < ni; ++i) {
                if (mask & (1u << i)) {
                    ns.c[ns.rank] = cs.c[i];
                    ns.p[ns.rank++] = weights[i];
                    g.ray += store[cs.c[i]].w * weights[i];
                } else {
                    g.freeList[g.nfree++] = cs.c[i];
                }
            }
            if (mask == 15) g.status = 1;
        } else {
            gjk2_remove(g, cs);
            break;
        }
        g.status = (++iterations < kGjk2MaxIter) ? g.status : 2;
    } while (g.status == 0);
    if (g.status == 0) g.distance = len(g.ray);
    else if (g.status == 1) g.distance = 0;
#if defined(__CUDA_ARCH__)
    EPA_T(2); EPA_N(3, 1); EPA_N(4, iterations);
#endif
    return g.status;
}

// GJK::EncloseOrigin (btGjkEpa2.cpp:334-410): complete the final simplex to a tetrahedron around the origin
template <class Sh>
RL_HD inline bool gjk2_enclose4(Gjk2& g, EpaSV* store, const Sh& sh) {
    (void)sh;
    const Gjk2Simplex& s = g.sx[g.current];
    return fabsf(det3(store[s.c[0]].w - store[s.c[3]].w, store[s.c[1]].w - store[s.c[3]].w, store[s.c[2]].w - store[s.c[3]].w)) > 0;
}
template <class Sh>
RL_HD inline bool gjk2_enclose3(Gjk2& g, EpaSV* store, const Sh& sh) {
    Gjk2Simplex& s = g.sx[g.current];
    const V3 n = cross(store[s.c[1]].w - store[s.c[0]].w, store[s.c[2]].w - store[s.c[0]].w);
    if (len2(n) > 0) {
        gjk2_append(g, store, sh, s, n);
        if (gjk2_enclose4(g, store, sh)) return true;
        gjk2_remove(g, s);
        gjk2_append(g, store, sh, s, -n);
        if (gjk2_enclose4(g, store, sh)) return true;
        gjk2_remove(g, s);
    }
    return false;
}
template <class Sh>
RL_HD inline bool gjk2_enclose2(Gjk2& g, EpaSV* store, const Sh& sh) {
    Gjk2Simplex& s = g.sx[g.current];
    const V3 d = store[s.c[1]].w - store[s.c[0]].w;
    for (int i = 0; i < 3; ++i) {
        V3 axis(0, 0, 0);
        axis[i] = 1;
        const V3 p = cross(d, axis);
        if (len2(p) > 0) {
            gjk2_append(g, store, sh, s, p);
            if (gjk2_enclose3(g, store, sh)) return true;
            gjk2_remove(g, s);
            gjk2_append(g, store, sh, s, -p);
            if (gjk2_enclose3(g, store, sh)) return true;
            gjk2_remove(g, s);
        }
    }
    return false;
}
template <class Sh>
RL_HD inline bool gjk2_enclose1(Gjk2& g, EpaSV* store, const Sh& sh) {
    Gjk2Simplex& s = g.sx[g.current];
    for (int i = 0; i < 3; ++i) {
        V3 axis(0, 0, 0);
        axis[i] = 1;
        gjk2_append(g, store, sh, s, axis);
        if (gjk2_enclose2(g, store, sh)) return true;
        gjk2_remove(g, s);
        gjk2_append(g, store, sh, s, -axis);
        if (gjk2_enclose2(g, store, sh)) return true;
        gjk2_remove(g, s);
    }
    return false;
}
template <class Sh>
RL_HD inline bool gjk2_enclose_origin(Gjk2& g, EpaSV* store, const Sh& sh) {
    switch (g.sx[g.current].rank) {
    case 1: return gjk2_enclose1(g, store, sh);
    case 2: return gjk2_enclose2(g, store, sh);
    case 3: return gjk2_enclose3(g, store, sh);
    case 4: return gjk2_enclose4(g, store, sh);
    }
    return false;
}

// ---- the polytope --------------------------------------------------------------------------------------------------
struct EpaList { int root, count; };
struct Epa {
    EpaWs ws;
    EpaList hull, stock;  // stock: the faces handed back (LIFO) on top of the never-used range [fresh, maxFaces)
    int fresh;
    bool overflow;        // ran out of room in a workspace smaller than the reference's
    int nextsv;
#if defined(RL_DEBUG_CONTACTS) && !defined(__CUDA_ARCH__)
    int dbgMaxFace = 0;
#endif
    int status;  // 0 valid, 1 touching, 2 degenerated, 3 non-convex, 4 invalid hull, 5 out of faces, 6 out of vertices, 7 accuracy reached, 8 fall-back, 9 failed
    V3 normal; float depth;
    Gjk2Simplex result;
};
enum { EPA_VALID = 0, EPA_DEGENERATED = 2, EPA_NONCONVEX = 3, EPA_INVALID_HULL = 4, EPA_OUT_OF_FACES = 5, EPA_OUT_OF_VERTICES = 6, EPA_ACCURACY_REACHED = 7,
       EPA_FALLBACK = 8, EPA_FAILED = 9 };

RL_HDI void epa_bind(EpaFace* fc, int fa, int ea, int fb, int eb) {
    fc[fa].e[ea] = (uint8_t)eb; fc[fa].f[ea] = (int16_t)fb;
    fc[fb].e[eb] = (uint8_t)ea; fc[fb].f[eb] = (int16_t)fa;
}
RL_HDI void epa_list_append(EpaFace* fc, EpaList& list, int face) {
    fc[face].l[0] = -1;
    fc[face].l[1] = (int16_t)list.root;
    if (list.root >= 0) fc[list.root].l[0] = (int16_t)face;
    list.root = face;
    ++list.count;
}
RL_HDI void epa_list_remove(EpaFace* fc, EpaList& list, int face) {
    if (fc[face].l[1] >= 0) fc[fc[face].l[1]].l[0] = fc[face].l[0];
    if (fc[face].l[0] >= 0) fc[fc[face].l[0]].l[1] = fc[face].l[1];
    if (face == list.root) list.root = fc[face].l[1];
    --list.count;
}
RL_HD inline void epa_init(Epa& e, const EpaWs& ws) {
    e.ws = ws;
    e.status = EPA_FAILED;
    e.normal = V3(0, 0, 0);
    e.depth = 0;
    e.nextsv = 0;
    e.hull.root = -1; e.hull.count = 0;
    // EPA::Initialize appends the faces to the stock from the back, so it hands them out in ascending index order; faces that
    // come back are pushed on top.  Kept as a LIFO of returned faces over a counter of never-used ones (no 256-entry set-up).
    e.stock.root = -1; e.stock.count = 0;
    e.fresh = 0;
    e.overflow = false;
}
// take the face EPA::newface would take from m_stock.root; -1: none left
RL_HDI int epa_stock_take(Epa& e) {
    EpaFace* fc = e.ws.fc;
    if (e.stock.root >= 0) { const int f = e.stock.root; epa_list_remove(fc, e.stock, f); return f; }
    if (e.fresh < e.ws.maxFaces) return e.fresh++;
    return -1;
}
// EPA::getedgedist: the origin projects outside edge a->b of the face -> distance to that edge or its end points
RL_HD inline bool epa_edge_dist(V3 n, V3 a, V3 b, float& dist) {
    const V3 ba = b - a;
    const V3 n_ab = cross(ba, n);
    const float a_dot_nab = dot(a, n_ab);
    if (a_dot_nab < 0) {
        const float ba_l2 = len2(ba);
        const float a_dot_ba = dot(a, ba);
        const float b_dot_ba = dot(b, ba);
        if (a_dot_ba > 0) dist = len(a);
        else if (b_dot_ba < 0) dist = len(b);
        else {
            const float a_dot_b = dot(a, b);
            dist = sqrtf(fmaxf_((len2(a) * len2(b) - a_dot_b * a_dot_b) / ba_l2, 0.f));
        }
        return true;
    }
    return false;
}
RL_HD inline int epa_newface(Epa& e, int a, int b, int c, bool forced) {
    EpaFace* fc = e.ws.fc;
    const EpaSV* sv = e.ws.sv;
    const int face = epa_stock_take(e);
    if (face >= 0) {
#if defined(RL_DEBUG_CONTACTS) && !defined(__CUDA_ARCH__)
        if (face > g_dbg_epa_maxface) g_dbg_epa_maxface = face;
        if (face > e.dbgMaxFace) e.dbgMaxFace = face;
#endif
        epa_list_append(fc, e.hull, face);
        EpaFace& f = fc[face];
        f.pass = 0;
        f.c[0] = (uint8_t)a; f.c[1] = (uint8_t)b; f.c[2] = (uint8_t)c;
        const V3 wa = sv[a].w, wb = sv[b].w, wc = sv[c].w;
        f.n = cross(wb - wa, wc - wa);
        const float l = len(f.n);
        const bool v = l > kEpaAccuracy;
        if (v) {
            if (!(epa_edge_dist(f.n, wa, wb, f.d) || epa_edge_dist(f.n, wb, wc, f.d) || epa_edge_dist(f.n, wc, wa, f.d)))
                f.d = dot(wa, f.n) / l;  // the origin projects into the face: plane distance
            f.n = f.n / l;
            if (forced || f.d >= -kEpaPlaneEps) return face;
            e.status = EPA_NONCONVEX;
        } else {
            e.status = EPA_DEGENERATED;
        }
        epa_list_remove(fc, e.hull, face);
        epa_list_append(fc, e.stock, face);
        return -1;
    }
    e.status = EPA_OUT_OF_FACES;
    if (e.ws.maxFaces < kEpaMaxFaces) e.overflow = true;
    return -1;
}
RL_HD inline int epa_findbest(const Epa& e) {
    const EpaFace* fc = e.ws.fc;
    int minf = e.hull.root;
    float mind = fc[minf].d * fc[minf].d;
    for (int f = fc[minf].l[1]; f >= 0; f = fc[f].l[1]) {
        const float sqd = fc[f].d * fc[f].d;
        if (sqd < mind) { minf = f; mind = sqd; }
    }
    return minf;
}
struct EpaHorizon { int cf, ff, nf; };
// EPA::expand: silhouette walk from (face f, edge e), depth first, edge e+1 before e+2, as an explicit stack.
// A frame is (face, edge, stage): stage 0 = entered, 1 = first child returned, 2 = second child returned.
RL_HD inline bool epa_expand(Epa& ep, unsigned pass, int w, int f0, int e0, EpaHorizon& hz) {
    EpaFace* fc = ep.ws.fc;
    const EpaSV* sv = ep.ws.sv;
    uint16_t* st = ep.ws.stack;
    int sp = 0;
    st[sp++] = (uint16_t)((f0 << 4) | (e0 << 2));
    bool ret = false;
    while (sp > 0) {
        const unsigned fr = st[sp - 1];
        const int f = (int)(fr >> 4), e = (int)((fr >> 2) & 3), stage = (int)(fr & 3);
        const int e1 = (e + 1) % 3, e2 = (e + 2) % 3;
        if (stage == 0) {
            if (fc[f].pass == pass) { ret = false; sp--; continue; }
            if ((dot(fc[f].n, sv[w].w) - fc[f].d) < -kEpaPlaneEps) {
                const int nf = epa_newface(ep, fc[f].c[e1], fc[f].c[e], w, false);
                if (nf >= 0) {
                    epa_bind(fc, nf, 0, f, e);
                    if (hz.cf >= 0) epa_bind(fc, hz.cf, 1, nf, 2);
                    else hz.ff = nf;
                    hz.cf = nf;
                    ++hz.nf;
                    ret = true;
                } else {
                    ret = false;
                }
                sp--;
                continue;
            }
            fc[f].pass = (uint8_t)pass;
            st[sp - 1] = (uint16_t)(fr | 1);
            st[sp++] = (uint16_t)((fc[f].f[e1] << 4) | (fc[f].e[e1] << 2));
            continue;
        }
        if (stage == 1) {
            if (!ret) { sp--; continue; }  // first child failed: the second is not visited
            st[sp - 1] = (uint16_t)((fr & ~3u) | 2);
            st[sp++] = (uint16_t)((fc[f].f[e2] << 4) | (fc[f].e[e2] << 2));
            continue;
        }
        // stage 2
        if (ret) {
            epa_list_remove(fc, ep.hull, f);
            epa_list_append(fc, ep.stock, f);
        }
        sp--;
    }
    return ret;
}

// EPA::Evaluate (btGjkEpa2.cpp:653-778)
template <class Sh>
RL_HD inline int epa_evaluate(Epa& ep, Gjk2& g, const Sh& sh, V3 guess) {
    EpaFace* fc = ep.ws.fc;
    EpaSV* sv = ep.ws.sv;
    Gjk2Simplex& sx = g.sx[g.current];
    if (sx.rank > 1 && gjk2_enclose_origin(g, sv, sh)) {
        ep.status = EPA_VALID;
        ep.nextsv = 0;
        if (det3(sv[sx.c[0]].w - sv[sx.c[3]].w, sv[sx.c[1]].w - sv[sx.c[3]].w, sv[sx.c[2]].w - sv[sx.c[3]].w) < 0) {
            uint8_t tc = sx.c[0]; sx.c[0] = sx.c[1]; sx.c[1] = tc;
            float tp = sx.p[0]; sx.p[0] = sx.p[1]; sx.p[1] = tp;
        }
        int tetra[4];
        tetra[0] = epa_newface(ep, sx.c[0], sx.c[1], sx.c[2], true);
        tetra[1] = epa_newface(ep, sx.c[1], sx.c[0], sx.c[3], true);
        tetra[2] = epa_newface(ep, sx.c[2], sx.c[1], sx.c[3], true);
        tetra[3] = epa_newface(ep, sx.c[0], sx.c[2], sx.c[3], true);
        if (ep.hull.count == 4) {
            int best = epa_findbest(ep);
            EpaFace outer = fc[best];
            unsigned pass = 0;
            epa_bind(fc, tetra[0], 0, tetra[1], 0);
            epa_bind(fc, tetra[0], 1, tetra[2], 0);
            epa_bind(fc, tetra[0], 2, tetra[3], 0);
            epa_bind(fc, tetra[1], 1, tetra[3], 2);
            epa_bind(fc, tetra[1], 2, tetra[2], 1);
            epa_bind(fc, tetra[2], 2, tetra[3], 1);
            ep.status = EPA_VALID;
            for (int iterations = 0; iterations < kEpaMaxIter; ++iterations) {
                if (ep.nextsv >= ep.ws.maxVerts && ep.ws.maxVerts < kEpaMaxVerts) { ep.overflow = true; break; }
                if (ep.nextsv < kEpaMaxVerts) {
                    EpaHorizon hz; hz.cf = -1; hz.ff = -1; hz.nf = 0;
                    const int w = 4 + ep.nextsv++;
                    bool valid = true;
#if defined(RL_DEBUG_CONTACTS) && !defined(__CUDA_ARCH__)
                    g_dbg_epa_iters++;
#endif
                    fc[best].pass = (uint8_t)(++pass);
                    gjk2_support(sh, fc[best].n, sv[w]);
                    const float wdist = dot(fc[best].n, sv[w].w) - fc[best].d;
                    if (wdist > kEpaAccuracy) {
                        for (int j = 0; j < 3 && valid; ++j) valid &= epa_expand(ep, pass, w, fc[best].f[j], fc[best].e[j], hz);
                        if (valid && hz.nf >= 3) {
                            epa_bind(fc, hz.cf, 1, hz.ff, 2);
                            epa_list_remove(fc, ep.hull, best);
                            epa_list_append(fc, ep.stock, best);
                            best = epa_findbest(ep);
                            outer = fc[best];
                        } else { ep.status = EPA_INVALID_HULL; break; }
                    } else { ep.status = EPA_ACCURACY_REACHED; break; }
                } else { ep.status = EPA_OUT_OF_VERTICES; break; }
            }
            const V3 projection = outer.n * outer.d;
            ep.normal = outer.n;
            ep.depth = outer.d;
            ep.result.rank = 3;
            ep.result.c[0] = outer.c[0]; ep.result.c[1] = outer.c[1]; ep.result.c[2] = outer.c[2];
            ep.result.p[0] = len(cross(sv[outer.c[1]].w - projection, sv[outer.c[2]].w - projection));
            ep.result.p[1] = len(cross(sv[outer.c[2]].w - projection, sv[outer.c[0]].w - projection));
            ep.result.p[2] = len(cross(sv[outer.c[0]].w - projection, sv[outer.c[1]].w - projection));
            const float sum = ep.result.p[0] + ep.result.p[1] + ep.result.p[2];
            ep.result.p[0] /= sum; ep.result.p[1] /= sum; ep.result.p[2] /= sum;
            return ep.status;
        }
    }
    ep.status = EPA_FALLBACK;
    ep.normal = -guess;
    const float nl = len(ep.normal);
    if (nl > 0) ep.normal = ep.normal / nl;
    else ep.normal = V3(1, 0, 0);
    ep.depth = 0;
    ep.result.rank = 1;
    ep.result.c[0] = sx.c[0];
    ep.result.p[0] = 1;
    return ep.status;
}

struct PenResult { V3 witnessA, witnessB, normal; };

// btGjkEpaSolver2::Penetration (btGjkEpa2.cpp:980-1027), margins on
template <class Sh>
RL_HD inline bool epa_penetration(const EpaWs& ws, Sh& sh, V3 guess, PenResult& r, bool& overflow) {
    sh.margins = true;
    Gjk2 g;
    const int gs = gjk2_evaluate(g, ws.sv, sh, -guess);
    if (gs == 1) {
        Epa ep;
        epa_init(ep, ws);
#if defined(__CUDA_ARCH__)
        EPA_T0();
#endif
        const int es = epa_evaluate(ep, g, sh, -guess);
#if defined(__CUDA_ARCH__)
        EPA_T(5); EPA_N(6, ep.nextsv);
#endif
        if (ep.overflow) { overflow = true; return false; }
#if defined(RL_DEBUG_CONTACTS) && !defined(__CUDA_ARCH__)
        if (ep.nextsv > g_dbg_epa_maxsv) g_dbg_epa_maxsv = ep.nextsv;
        { int b = ep.dbgMaxFace < 8 ? 0 : ep.dbgMaxFace < 12 ? 1 : ep.dbgMaxFace < 16 ? 2 : ep.dbgMaxFace < 24 ? 3 : ep.dbgMaxFace < 32 ? 4 : ep.dbgMaxFace < 48 ? 5 : ep.dbgMaxFace < 64 ? 6 : 7; g_dbg_epa_hist[b]++; }
#endif
        if (es != EPA_FAILED) {
            V3 w0(0, 0, 0);
            for (int i = 0; i < ep.result.rank; ++i) w0 += sh.support0(ws.sv[ep.result.c[i]].d) * ep.result.p[i];
            r.witnessA = sh.to_world(w0);
            r.witnessB = sh.to_world(w0 - ep.normal * ep.depth);
            r.normal = -ep.normal;
            return true;
        }
    }
    return false;
}
// btGjkEpaSolver2::Distance (btGjkEpa2.cpp:944-976), margins off
template <class Sh>
RL_HD inline bool epa_distance(const EpaWs& ws, Sh& sh, V3 guess, PenResult& r) {
    sh.margins = false;
    Gjk2 g;
    const int gs = gjk2_evaluate(g, ws.sv, sh, guess);
    if (gs == 0) {
        V3 w0(0, 0, 0), w1(0, 0, 0);
        const Gjk2Simplex& sx = g.sx[g.current];
        for (int i = 0; i < sx.rank; ++i) {
            const float p = sx.p[i];
            w0 += sh.support0(ws.sv[sx.c[i]].d) * p;
            w1 += sh.support1(-ws.sv[sx.c[i]].d) * p;
        }
        r.witnessA = sh.to_world(w0);
        r.witnessB = sh.to_world(w1);
        r.normal = w0 - w1;
        const float distance = len(r.normal);
        r.normal = r.normal / (distance > kGjk2MinDist ? distance : 1.f);
        return true;
    }
    return false;
}

// btGjkEpaPenetrationDepthSolver::calcPenDepth: -> 1 penetration found, 0 separated (witnesses valid, v set), -1 nothing (v = 0),
// -2 the (small) workspace ran out of room: repeat in a larger one
template <class Sh>
RL_HD inline int calc_pen_depth(const EpaWs& ws, Sh& sh, V3 originA, V3 originB, PenResult& r) {
#if defined(RL_DEBUG_CONTACTS) && !defined(__CUDA_ARCH__)
    g_dbg_epa_calls++;
#endif
    const V3 guesses[9] = {safe_normalized(originB - originA), safe_normalized(originA - originB), V3(0, 0, 1), V3(0, 1, 0), V3(1, 0, 0),
                           V3(1, 1, 0), V3(1, 1, 1), V3(0, 1, 1), V3(1, 0, 1)};
    for (int i = 0; i < 9; i++) {
        bool overflow = false;
#if defined(__CUDA_ARCH__)
        EPA_N(7, 1);
#endif
        if (epa_penetration(ws, sh, guesses[i], r, overflow)) return 1;
        if (overflow) return -2;
        if (epa_distance(ws, sh, guesses[i], r)) return 0;
    }
    r.witnessA = r.witnessB = r.normal = V3(0, 0, 0);
    return -1;
}

}  // namespace rl
