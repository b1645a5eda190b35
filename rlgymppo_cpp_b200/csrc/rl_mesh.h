// rl_mesh.h — the static world: soccar's 4 planes + the arena collision meshes with a BVH.
//
// The mesh set is read-only and shared by every arena (reference: one
// btBvhTriangleMeshShape per .cmf shared across Arenas, R/RocketSim.cpp:70-212).  It
// is a few hundred KB at most and stays resident in the 126 MB L2.
//
// BVH: built on the host with the same splitting rule as the reference's
// btQuantizedBvh::buildTree (B/BulletCollision/BroadphaseCollision/btQuantizedBvh.cpp:117-300:
// axis of largest centre variance, partition around the mean, balanced fallback) so
// that leaves are visited in the same order; boxes are kept as fp32 (the reference
// quantises to 16 bit, which only makes its culling more conservative).  Nodes are
// laid out depth-first with subtree sizes for stackless traversal.
#pragma once
#include "rl_state.h"

namespace rl {

struct BvhNode {  // 32 bytes
    float mn[3];
    int32_t tri;     // >= 0: leaf -> triangle index (global); -1: internal
    float mx[3];
    int32_t escape;  // internal: number of nodes in this subtree (incl. itself)
};

struct Tri {
    V3 v0, v1, v2;
};

constexpr int kMaxMeshes = 16;

struct MeshSet {
    int32_t numMeshes;
    int32_t nodeStart[kMaxMeshes];    // first BVH node of mesh m
    int32_t nodeCount[kMaxMeshes];
    // AABB-query traversal order: the reference walks "subtree headers"
    // (btQuantizedBvh::walkStacklessQuantizedTreeCacheFriendly); ranges [hdrStart[m], hdrStart[m+1])
    int32_t hdrStart[kMaxMeshes + 1];
    const int32_t* hdrRoot;   // node index (global)
    const int32_t* hdrSize;   // nodes in subtree
    const BvhNode* nodes;
    const Tri* tris;
    // per triangle: internal-edge info (btTriangleInfo: flags + 3 edge angles)
    const int32_t* triFlags;
    const float* triEdgeAngles;  // [T][3]
    int32_t numTris, numNodes, numHdrs;
    // "free box": an axis-aligned region that no (quantised, padded) leaf box of any mesh touches.  A query whose AABB
    // lies inside it cannot reach a triangle, so the BVH walks are skipped with the same result (the arena meshes hug
    // the perimeter; most bodies are in the open field).  Empty (mn > mx) when a mesh covers the centre.
    V3 freeMn, freeMx;
    // Leaf grid: the arena is static, so the set of leaf boxes a SMALL query box can overlap is precomputed per cell of
    // a uniform grid: cell (ix, iy, iz) lists, in (mesh, depth-first leaf) order, every leaf whose box overlaps the cell
    // grown by gridMargin on each side.  A query box whose half extents are <= gridMargin lies inside the grown cell of
    // its centre, so scanning that one list with the same per-leaf overlap test returns exactly the leaves — in
    // exactly the order — of the stackless BVH walks, as a short flat scan of independent loads instead of a chain of
    // dependent node fetches (host_build_leaf_grid).  Bigger / out-of-grid queries fall back to the walks.
    V3 gridOrigin;
    float gridCell, gridInvCell, gridMargin;
    int32_t gridNx, gridNy, gridNz;
    const uint32_t* gridRange;  // per cell: first entry of gridList (24 bits) << 8 | min(count, 255); 255 = "walk the BVH"
    const int32_t* gridList;    // leaf node index (global) | mesh << 24
};
RL_HDI bool inside_free_box(const MeshSet& ms, V3 mn, V3 mx) {
#ifdef RL_NO_FREEBOX
    return false;
#endif
    return mn.x > ms.freeMn.x && mx.x < ms.freeMx.x && mn.y > ms.freeMn.y && mx.y < ms.freeMx.y && mn.z > ms.freeMn.z && mx.z < ms.freeMx.z;
}

// the 4 soccar planes (R/Sim/Arena/Arena.cpp:1060-1101), Bullet units; point on plane + normal
struct PlaneDef { V3 n; V3 origin; };
RL_HDI PlaneDef world_plane(int i) {
    const float ex = C::ARENA_EXTENT_X, h = C::ARENA_HEIGHT;
    switch (i) {
    case 0: return PlaneDef{V3(0, 0, 1), V3(0, 0, 0)};
    case 1: return PlaneDef{V3(0, 0, -1), V3(0.f * UU2BT, 0.f * UU2BT, h * UU2BT)};
    case 2: return PlaneDef{V3(1, 0, 0), V3(-ex * UU2BT, 0.f * UU2BT, (h / 2) * UU2BT)};
    default: return PlaneDef{V3(-1, 0, 0), V3(ex * UU2BT, 0.f * UU2BT, (h / 2) * UU2BT)};
    }
}

RL_HDI bool aabb_overlap(const float* amn, const float* amx, V3 bmn, V3 bmx) {
    return !(amn[0] > bmx.x || amx[0] < bmn.x || amn[1] > bmx.y || amx[1] < bmn.y || amn[2] > bmx.z || amx[2] < bmn.z);
}

// btRayAabb2-style slab test on [0, maxFrac]
RL_HDI bool ray_aabb(V3 from, V3 invDir, const float* mn, const float* mx, float maxFrac) {
    float tmin = 0.f, tmax = maxFrac;
    for (int a = 0; a < 3; a++) {
        float t1 = (mn[a] - from[a]) * invDir[a];
        float t2 = (mx[a] - from[a]) * invDir[a];
        float lo = fminf_(t1, t2), hi = fmaxf_(t1, t2);
        tmin = fmaxf_(tmin, lo);
        tmax = fminf_(tmax, hi);
    }
    return tmin <= tmax;
}

struct RayHit {
    float frac;
    V3 normal;
    int body;  // -2 none, -1 static world, >=0 dynamic body index (0 ball, 1+c car c)
};

// btTriangleRaycastCallback::processTriangle (B/BulletCollision/NarrowPhaseCollision/btRaycastCallback.cpp:34-111), flags = 0
RL_HDI void ray_triangle(V3 from, V3 to, V3 vert0, V3 vert1, V3 vert2, RayHit& hit) {
    V3 v10 = vert1 - vert0, v20 = vert2 - vert0;
    V3 n = cross(v10, v20);
    float dist = dot(vert0, n);
    float dist_a = dot(n, from) - dist;
    float dist_b = dot(n, to) - dist;
    if (dist_a * dist_b >= 0.f) return;
    float proj_length = dist_a - dist_b;
    float distance = dist_a / proj_length;
    if (distance < hit.frac) {
        float edge_tol = len2(n) * -0.0001f;
        V3 point = from + (to - from) * distance;  // setInterpolate3
        V3 v0p = vert0 - point, v1p = vert1 - point;
        if (dot(cross(v0p, v1p), n) >= edge_tol) {
            V3 v2p = vert2 - point;
            if (dot(cross(v1p, v2p), n) >= edge_tol && dot(cross(v2p, v0p), n) >= edge_tol) {
                V3 nn = normalized(n);
                hit.frac = distance;
                hit.normal = dist_a <= 0.f ? -nn : nn;
                hit.body = -1;
            }
        }
    }
}

// ray vs infinite plane == btStaticPlaneShape::processAllTriangles fed to the ray-triangle test
// (the two generated triangles always contain the intersection point)
RL_HDI void ray_plane(V3 from, V3 to, const PlaneDef& p, RayHit& hit) {
    float dist_a = dot(p.n, from - p.origin);
    float dist_b = dot(p.n, to - p.origin);
    if (dist_a * dist_b >= 0.f) return;
    float distance = dist_a / (dist_a - dist_b);
    if (distance < hit.frac) {
        hit.frac = distance;
        hit.normal = dist_a <= 0.f ? -p.n : p.n;
        hit.body = -1;
    }
}

RL_HD inline void ray_meshes(V3 from, V3 to, const MeshSet& ms, RayHit& hit) {
    if (inside_free_box(ms, vmin(from, to), vmax(from, to))) return;
    V3 d = to - from;
    V3 inv(1.f / d.x, 1.f / d.y, 1.f / d.z);
    for (int m = 0; m < ms.numMeshes; m++) {
        int i = ms.nodeStart[m], end = ms.nodeStart[m] + ms.nodeCount[m];
        while (i < end) {
            const BvhNode& nd = ms.nodes[i];
            // pad the box a little: the reference's quantised boxes are conservative too
            float mn[3] = {nd.mn[0] - 0.01f, nd.mn[1] - 0.01f, nd.mn[2] - 0.01f};
            float mx[3] = {nd.mx[0] + 0.01f, nd.mx[1] + 0.01f, nd.mx[2] + 0.01f};
            bool ov = ray_aabb(from, inv, mn, mx, 1.0f);
            if (nd.tri >= 0) {
                if (ov) { const Tri& t = ms.tris[nd.tri]; ray_triangle(from, to, t.v0, t.v1, t.v2, hit); }
                i++;
            } else {
                i += ov ? 1 : nd.escape;
            }
        }
    }
}

// ---- one BVH query per car per tick -----------------------------------------------------------------------------
// The four wheel rays and the hitbox of a car live within a metre of each other, so instead of five BVH walks per car
// per tick (4 x ray_meshes + box_meshes) the car collects ONE candidate list for the union AABB (in the meshes' DFS
// leaf order, i.e. the order the reference visits them) and runs the exact per-triangle tests on that list.  A triangle
// outside a ray's / the hitbox's own box cannot pass its exact test, so results are unchanged.  Overflowing lists fall
// back to the direct walks.
constexpr int kMaxCands = 24;
struct MeshCands {
    int32_t n;        // -1: overflow -> callers use the direct BVH walks
    int32_t node[kMaxCands];  // global BVH leaf-node index | mesh << 24
};
// Leaf-grid lookup for the query box [mn, mx]: true + the list range when the grid answers it.
RL_HDI bool grid_lookup(const MeshSet& ms, V3 mn, V3 mx, int& first, int& count) {
#ifdef RL_NO_LEAF_GRID
    return false;
#endif
    if (!ms.gridRange) return false;
    V3 h = (mx - mn) * 0.5f, c = (mn + mx) * 0.5f;
    if (!(fmaxf_(fmaxf_(h.x, h.y), h.z) <= ms.gridMargin)) return false;
    V3 r = (c - ms.gridOrigin) * ms.gridInvCell;
    if (!(r.x >= 0.f && r.y >= 0.f && r.z >= 0.f)) return false;
    int ix = (int)r.x, iy = (int)r.y, iz = (int)r.z;
    if (ix >= ms.gridNx || iy >= ms.gridNy || iz >= ms.gridNz) return false;
    uint32_t e = ms.gridRange[(iz * ms.gridNy + iy) * ms.gridNx + ix];
    count = (int)(e & 255u);
    first = (int)(e >> 8);
    return count != 255;
}

RL_HD inline void collect_candidates(const MeshSet& ms, V3 mn, V3 mx, MeshCands& out) {
    out.n = 0;
    if (inside_free_box(ms, mn, mx)) return;
    int first, count;
    if (grid_lookup(ms, mn, mx, first, count)) {
        for (int j = 0; j < count; j++) {
            int e = ms.gridList[first + j];
            const BvhNode& nd = ms.nodes[e & 0xffffff];
            if (aabb_overlap(nd.mn, nd.mx, mn, mx)) {
                if (out.n >= kMaxCands) { out.n = -1; return; }
                out.node[out.n++] = e;
            }
        }
        return;
    }
    for (int m = 0; m < ms.numMeshes; m++) {
        int i = ms.nodeStart[m], end = ms.nodeStart[m] + ms.nodeCount[m];
        while (i < end) {
            const BvhNode& nd = ms.nodes[i];
            bool ov = aabb_overlap(nd.mn, nd.mx, mn, mx);
            if (nd.tri >= 0) {
                if (ov) {
                    if (out.n >= kMaxCands) { out.n = -1; return; }
                    out.node[out.n++] = i | (m << 24);
                }
                i++;
            } else {
                i += ov ? 1 : nd.escape;
            }
        }
    }
}
// One triangle against the four wheel rays of a car, independent of any earlier hit: d[i] = hit fraction along ray i
// (2 = no hit), nn = unit triangle normal, bit i of neg = ray i starts behind the triangle (hit normal = -nn).  The same
// arithmetic as ray_triangle; the caller applies ray_triangle's `distance < hit.frac` gate in candidate order
// (ray_tri4_apply), so scanning the candidates triangle-major gives the same hits as four ray-major scans.
struct TriRays { float d[4]; V3 nn; uint32_t neg; };
RL_HDI void ray_tri4(const V3* from, const V3* to, V3 vert0, V3 vert1, V3 vert2, TriRays& out) {
    V3 v10 = vert1 - vert0, v20 = vert2 - vert0;
    V3 n = cross(v10, v20);
    float dist = dot(vert0, n);
    float edge_tol = len2(n) * -0.0001f;
    out.neg = 0; out.nn = V3();
    bool any = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        out.d[i] = 2.f;
        float dist_a = dot(n, from[i]) - dist;
        float dist_b = dot(n, to[i]) - dist;
        if (dist_a * dist_b >= 0.f) continue;
        float proj_length = dist_a - dist_b;
        float distance = dist_a / proj_length;
        V3 point = from[i] + (to[i] - from[i]) * distance;  // setInterpolate3
        V3 v0p = vert0 - point, v1p = vert1 - point;
        if (dot(cross(v0p, v1p), n) >= edge_tol) {
            V3 v2p = vert2 - point;
            if (dot(cross(v1p, v2p), n) >= edge_tol && dot(cross(v2p, v0p), n) >= edge_tol) {
                out.d[i] = distance;
                if (dist_a <= 0.f) out.neg |= 1u << i;
                any = true;
            }
        }
    }
    if (any) out.nn = normalized(n);
}
RL_HDI void ray_tri4_apply(const float* d, V3 nn, uint32_t neg, RayHit* hit) {
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (d[i] < hit[i].frac) {
            hit[i].frac = d[i];
            hit[i].normal = ((neg >> i) & 1u) ? -nn : nn;
            hit[i].body = -1;
        }
}

// ray vs sphere (what the sub-simplex convex cast of a point against btSphereShape converges to)
RL_HDI void ray_sphere(V3 from, V3 to, V3 c, float r, int body, RayHit& hit) {
    V3 d = to - from, m = from - c;
    float a = dot(d, d), b = dot(m, d), cc = dot(m, m) - r * r;
    if (cc > 0.f && b > 0.f) return;
    float disc = b * b - a * cc;
    if (disc < 0.f) return;
    float t = (-b - sqrtf(disc)) / a;
    if (t < 0.f) t = 0.f;
    if (t < hit.frac) {
        V3 p = from + d * t;
        hit.frac = t; hit.normal = safe_normalized(p - c); hit.body = body;
    }
}

// ray vs oriented box (full extents; the cast uses localGetSupportingVertex = core + margin)
RL_HDI void ray_obb(V3 from, V3 to, V3 center, const M3& rot, V3 half, int body, RayHit& hit) {
    V3 lf = tmul(from - center, rot), lt = tmul(to - center, rot);
    V3 d = lt - lf;
    float tmin = 0.f, tmax = 1.f;
    int axis = -1; float sign = 0.f;
    for (int a = 0; a < 3; a++) {
        if (fabsf(d[a]) < 1e-12f) {
            if (lf[a] < -half[a] || lf[a] > half[a]) return;
        } else {
            float inv = 1.f / d[a];
            float t1 = (-half[a] - lf[a]) * inv, t2 = (half[a] - lf[a]) * inv;
            float s = -1.f;
            if (t1 > t2) { float t = t1; t1 = t2; t2 = t; s = 1.f; }
            if (t1 > tmin) { tmin = t1; axis = a; sign = s; }
            if (t2 < tmax) tmax = t2;
            if (tmin > tmax) return;
        }
    }
    if (axis < 0) return;  // started inside
    if (tmin < hit.frac) {
        V3 ln(0, 0, 0); ln[axis] = sign;
        hit.frac = tmin; hit.normal = rot * ln; hit.body = body;
    }
}

}  // namespace rl
