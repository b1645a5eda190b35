// rl_host_build.h — HOST-only builders: .cmf parsing + BVH construction, lookup tables, SimCfg.
//
// BVH build follows btQuantizedBvh::buildTree / calcSplittingAxis / sortAndCalcSplittingIndex
// (B/BulletCollision/BroadphaseCollision/btQuantizedBvh.cpp:117-300) and btOptimizedBvh::build's
// leaf boxes (min extent padding 0.002), so leaves end up in the reference's order; the
// "subtree header" list reproduces the order walkStacklessQuantizedTreeCacheFriendly visits them
// (subtrees <= 2048 bytes of 16-byte quantised nodes = 128 nodes).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "rl_gym.h"
#include "rl_mesh.h"

namespace rl {

struct HostMeshSet {
    std::vector<Tri> tris;
    std::vector<BvhNode> nodes;
    std::vector<int32_t> hdrRoot, hdrSize, triFlags;
    std::vector<float> triEdgeAngles;
    std::vector<uint32_t> gridRange;
    std::vector<int32_t> gridList;
    MeshSet meta;  // pointers unset
};

namespace detail {
struct Leaf { V3 mn, mx; int tri; };

struct BvhBuilder {
    std::vector<Leaf> leaves;
    std::vector<BvhNode>& nodes;
    std::vector<int32_t>& hdrRoot;
    std::vector<int32_t>& hdrSize;
    int nodeBase;
    explicit BvhBuilder(std::vector<BvhNode>& n, std::vector<int32_t>& hr, std::vector<int32_t>& hs) : nodes(n), hdrRoot(hr), hdrSize(hs), nodeBase(0) {}

    static V3 center(const Leaf& l) { return (l.mx + l.mn) * 0.5f; }

    int calcSplittingAxis(int s, int e) {
        V3 means(0, 0, 0), variance(0, 0, 0);
        int n = e - s;
        for (int i = s; i < e; i++) means += center(leaves[i]);
        means *= (1.f / (float)n);
        for (int i = s; i < e; i++) { V3 d = center(leaves[i]) - means; variance += d * d; }
        variance *= (1.f / ((float)n - 1));
        return variance.x < variance.y ? (variance.y < variance.z ? 2 : 1) : (variance.x < variance.z ? 2 : 0);
    }
    int sortAndCalcSplittingIndex(int s, int e, int axis) {
        int splitIndex = s, n = e - s;
        V3 means(0, 0, 0);
        for (int i = s; i < e; i++) means += center(leaves[i]);
        means *= (1.f / (float)n);
        float splitValue = means[axis];
        for (int i = s; i < e; i++) {
            if (center(leaves[i])[axis] > splitValue) { std::swap(leaves[i], leaves[splitIndex]); splitIndex++; }
        }
        int range = n / 3;
        bool unbalanced = (splitIndex <= (s + range)) || (splitIndex >= (e - 1 - range));
        if (unbalanced) splitIndex = s + (n >> 1);
        return splitIndex;
    }
    static int subtreeSize(const BvhNode& n) { return n.tri >= 0 ? 1 : n.escape; }
    void build(int s, int e) {
        int cur = (int)nodes.size();
        if (e - s == 1) {
            BvhNode nd;
            for (int a = 0; a < 3; a++) { nd.mn[a] = leaves[s].mn[a]; nd.mx[a] = leaves[s].mx[a]; }
            nd.tri = leaves[s].tri; nd.escape = 1;
            nodes.push_back(nd);
            return;
        }
        int axis = calcSplittingAxis(s, e);
        int split = sortAndCalcSplittingIndex(s, e, axis);
        BvhNode nd;
        V3 mn(1e30f, 1e30f, 1e30f), mx(-1e30f, -1e30f, -1e30f);
        for (int i = s; i < e; i++) { mn = vmin(mn, leaves[i].mn); mx = vmax(mx, leaves[i].mx); }
        for (int a = 0; a < 3; a++) { nd.mn[a] = mn[a]; nd.mx[a] = mx[a]; }
        nd.tri = -1; nd.escape = 0;
        nodes.push_back(nd);
        int left = (int)nodes.size();
        build(s, split);
        int right = (int)nodes.size();
        build(split, e);
        int escape = (int)nodes.size() - cur;
        nodes[cur].escape = escape;
        const int MAX_NODES = 2048 / 16;
        if (escape > MAX_NODES) {  // updateSubtreeHeaders
            int ls = subtreeSize(nodes[left]), rs = subtreeSize(nodes[right]);
            if (ls <= MAX_NODES) { hdrRoot.push_back(left); hdrSize.push_back(ls); }
            if (rs <= MAX_NODES) { hdrRoot.push_back(right); hdrSize.push_back(rs); }
        }
    }
};
}  // namespace detail

// btGenerateInternalEdgeInfo / btConnectivityProcessor (B/BulletCollision/CollisionDispatch/btInternalEdgeUtility.cpp:52-352):
// for every triangle A, every triangle B of the same mesh overlapping A's AABB and sharing exactly two vertices
// contributes the signed dihedral angle of the shared edge + convex / swap flags.
inline void host_build_edge_info(HostMeshSet& out) {
    const MeshSet& ms = out.meta;
    const float equalVertexThreshold = 0.0001f * 0.0001f, planarEpsilon = 0.0001f;
    const float PI = 3.1415926535897932384626433832795029f;
    auto quatRot = [](V3 axis, float angle, V3 v) {
        float d = len(axis); float s = sinf(angle * 0.5f) / d;
        Quat q(axis.x * s, axis.y * s, axis.z * s, cosf(angle * 0.5f));
        Quat t(q.w * v.x + q.y * v.z - q.z * v.y, q.w * v.y + q.z * v.x - q.x * v.z, q.w * v.z + q.x * v.y - q.y * v.x, -q.x * v.x - q.y * v.y - q.z * v.z);
        Quat r = t * Quat(-q.x, -q.y, -q.z, q.w);
        return V3(r.x, r.y, r.z);
    };
    for (int m = 0; m < ms.numMeshes; m++) {
        int nodeEnd = ms.nodeStart[m] + ms.nodeCount[m];
        // triangles of this mesh = leaves of its BVH
        std::vector<int> tris;
        for (int i = ms.nodeStart[m]; i < nodeEnd; i++) if (out.nodes[i].tri >= 0) tris.push_back(out.nodes[i].tri);
        for (int ta : tris) {
            const Tri& A = out.tris[ta];
            V3 va[3] = {A.v0, A.v1, A.v2};
            V3 amn = vmin(vmin(A.v0, A.v1), A.v2), amx = vmax(vmax(A.v0, A.v1), A.v2);
            int i = ms.nodeStart[m];
            while (i < nodeEnd) {
                const BvhNode& nd = out.nodes[i];
                float mn[3] = {nd.mn[0] - 0.01f, nd.mn[1] - 0.01f, nd.mn[2] - 0.01f}, mx[3] = {nd.mx[0] + 0.01f, nd.mx[1] + 0.01f, nd.mx[2] + 0.01f};
                bool ov = aabb_overlap(mn, mx, amn, amx);
                if (nd.tri < 0) { i += ov ? 1 : nd.escape; continue; }
                i++;
                if (!ov || nd.tri == ta) continue;
                const Tri& B = out.tris[nd.tri];
                V3 vb[3] = {B.v0, B.v1, B.v2};
                if (len2(cross(vb[1] - vb[0], vb[2] - vb[0])) < equalVertexThreshold) continue;
                if (len2(cross(va[1] - va[0], va[2] - va[0])) < equalVertexThreshold) continue;
                int numshared = 0, sa[3] = {-1, -1, -1}, sbv[3] = {-1, -1, -1};
                bool degenerate = false;
                for (int p = 0; p < 3 && !degenerate; p++) {
                    for (int q = 0; q < 3; q++) {
                        if (len2(va[p] - vb[q]) < equalVertexThreshold) {
                            sa[numshared] = p; sbv[numshared] = q; numshared++;
                            if (numshared >= 3) { degenerate = true; break; }
                        }
                    }
                }
                if (degenerate || numshared != 2) continue;
                if (sa[0] == 0 && sa[1] == 2) { sa[0] = 2; sa[1] = 0; int t = sbv[1]; sbv[1] = sbv[0]; sbv[0] = t; }
                int sumvertsA = sa[0] + sa[1];
                int otherIndexA = 3 - sumvertsA;
                V3 edge = normalized(va[sa[1]] - va[sa[0]]);
                int otherIndexB = 3 - (sbv[0] + sbv[1]);
                V3 tb0 = vb[sbv[1]], tb1 = vb[sbv[0]], tb2 = vb[otherIndexB];
                V3 normalA = normalized(cross(va[1] - va[0], va[2] - va[0]));
                V3 normalB = normalized(cross(tb1 - tb0, tb2 - tb0));
                V3 edgeCrossA = normalized(cross(edge, normalA));
                if (dot(edgeCrossA, va[otherIndexA] - va[sa[0]]) < 0) edgeCrossA = edgeCrossA * -1.f;
                V3 edgeCrossB = normalized(cross(edge, normalB));
                if (dot(edgeCrossB, vb[otherIndexB] - vb[sbv[0]]) < 0) edgeCrossB = edgeCrossB * -1.f;
                float angle2 = 0, ang4 = 0, correctedAngle = 0;
                bool isConvex = false;
                V3 calculatedEdge = cross(edgeCrossA, edgeCrossB);
                if (len2(calculatedEdge) < planarEpsilon) {
                    angle2 = 0; ang4 = 0;
                } else {
                    calculatedEdge = normalized(calculatedEdge);
                    V3 calculatedNormalA = normalized(cross(calculatedEdge, edgeCrossA));
                    angle2 = atan2f(dot(edgeCrossB, calculatedNormalA), dot(edgeCrossB, edgeCrossA));
                    ang4 = PI - angle2;
                    isConvex = dot(normalA, edgeCrossB) < 0.f;
                    correctedAngle = isConvex ? ang4 : -ang4;
                }
                int e; V3 eAxis;
                if (sumvertsA == 1) { e = 0; eAxis = va[0] - va[1]; }
                else if (sumvertsA == 2) { e = 2; eAxis = va[2] - va[0]; }
                else { e = 1; eAxis = va[1] - va[2]; }
                V3 computedNormalB = quatRot(eAxis, -correctedAngle, normalA);
                if (dot(computedNormalB, normalB) < 0) out.triFlags[ta] |= (8 << e);
                out.triEdgeAngles[ta * 3 + e] = -correctedAngle;
                if (isConvex) out.triFlags[ta] |= (1 << e);
            }
        }
    }
}

// Leaf grid (rl_mesh.h MeshSet::gridRange/gridList): per cell, the leaves (mesh-major, depth-first order = the order of
// the stackless walks) whose box overlaps the cell grown by margin + slack.  The slack absorbs the rounding of the
// query's centre / half extents so that "half extents <= margin" really implies "inside the grown cell".
inline void host_build_leaf_grid(HostMeshSet& out) {
    MeshSet& ms = out.meta;
    ms.gridRange = nullptr; ms.gridList = nullptr;
    out.gridRange.clear(); out.gridList.clear();
    if (out.nodes.empty()) return;
    const float CELL = 5.f, MARGIN = 2.6f, SLACK = 0.05f;  // Bullet units (250 uu cells; ball box 1.905, car + wheel rays < 2.6)
    V3 lo(1e30f, 1e30f, 1e30f), hi(-1e30f, -1e30f, -1e30f);
    for (const BvhNode& nd : out.nodes) if (nd.tri >= 0) {
        lo = vmin(lo, V3(nd.mn[0], nd.mn[1], nd.mn[2])); hi = vmax(hi, V3(nd.mx[0], nd.mx[1], nd.mx[2]));
    }
    // bodies can only be where the meshes / the four planes let them; cover the leaf bounds plus one cell
    lo = lo - V3(CELL, CELL, CELL); hi = hi + V3(CELL, CELL, CELL);
    int nx = (int)std::ceil((hi.x - lo.x) / CELL), ny = (int)std::ceil((hi.y - lo.y) / CELL), nz = (int)std::ceil((hi.z - lo.z) / CELL);
    if (nx < 1 || ny < 1 || nz < 1 || (int64_t)nx * ny * nz > (int64_t)4 << 20) return;  // degenerate / absurd extents: BVH walks only
    ms.gridOrigin = lo; ms.gridCell = CELL; ms.gridInvCell = 1.f / CELL; ms.gridMargin = MARGIN;
    ms.gridNx = nx; ms.gridNy = ny; ms.gridNz = nz;
    std::vector<std::vector<int32_t>> cells((size_t)nx * ny * nz);
    const float G = MARGIN + SLACK;
    for (int m = 0; m < ms.numMeshes; m++) {
        for (int i = ms.nodeStart[m]; i < ms.nodeStart[m] + ms.nodeCount[m]; i++) {
            const BvhNode& nd = out.nodes[i];
            if (nd.tri < 0) continue;
            int c0[3], c1[3];
            const int dims[3] = {nx, ny, nz};
            for (int a = 0; a < 3; a++) {
                // cell k grown: [lo + k*CELL - G, lo + (k+1)*CELL + G] overlaps [mn, mx]
                c0[a] = (int)std::floor((nd.mn[a] - G - lo[a]) / CELL) - 1;
                c1[a] = (int)std::floor((nd.mx[a] + G - lo[a]) / CELL) + 1;
                c0[a] = std::max(c0[a], 0); c1[a] = std::min(c1[a], dims[a] - 1);
            }
            for (int iz = c0[2]; iz <= c1[2]; iz++) for (int iy = c0[1]; iy <= c1[1]; iy++) for (int ix = c0[0]; ix <= c1[0]; ix++) {
                float cmn[3] = {lo.x + ix * CELL - G, lo.y + iy * CELL - G, lo.z + iz * CELL - G};
                float cmx[3] = {lo.x + (ix + 1) * CELL + G, lo.y + (iy + 1) * CELL + G, lo.z + (iz + 1) * CELL + G};
                bool ov = !(nd.mn[0] > cmx[0] || nd.mx[0] < cmn[0] || nd.mn[1] > cmx[1] || nd.mx[1] < cmn[1] || nd.mn[2] > cmx[2] || nd.mx[2] < cmn[2]);
                if (ov) cells[((size_t)iz * ny + iy) * nx + ix].push_back(i | (m << 24));
            }
        }
    }
    out.gridRange.resize(cells.size());
    for (size_t c = 0; c < cells.size(); c++) {
        size_t first = out.gridList.size();
        if (cells[c].size() >= 255 || first >= ((size_t)1 << 24)) { out.gridRange[c] = 255u; continue; }  // too long: walk the BVH there
        out.gridRange[c] = (uint32_t)(first << 8) | (uint32_t)cells[c].size();
        out.gridList.insert(out.gridList.end(), cells[c].begin(), cells[c].end());
    }
    if (out.gridList.empty()) out.gridList.push_back(0);
}

inline void host_build_meshes(const void* const* blobs, const size_t* sizes, int n, HostMeshSet& out) {
    if (n > kMaxMeshes) throw std::runtime_error("too many collision meshes");
    out = HostMeshSet();
    MeshSet& ms = out.meta;
    std::memset(&ms, 0, sizeof(ms));
    ms.numMeshes = n;
    for (int m = 0; m < n; m++) {
        const uint8_t* b = (const uint8_t*)blobs[m];
        if (sizes[m] < 8) throw std::runtime_error("collision mesh blob too small");
        int32_t nt, nv;
        std::memcpy(&nt, b, 4); std::memcpy(&nv, b + 4, 4);
        if (nt <= 0 || nv <= 0 || nt > 1000000 || nv > 1000000) throw std::runtime_error("bad triangle/vertex count in collision mesh");
        size_t need = 8 + (size_t)nt * 12 + (size_t)nv * 12;
        if (sizes[m] < need) throw std::runtime_error("collision mesh blob truncated");
        const int32_t* idx = (const int32_t*)(b + 8);
        const float* vtx = (const float*)(b + 8 + (size_t)nt * 12);
        int triBase = (int)out.tris.size();
        detail::BvhBuilder bb(out.nodes, out.hdrRoot, out.hdrSize);
        std::vector<Tri> local;
        V3 meshMn(1e30f, 1e30f, 1e30f), meshMx(-1e30f, -1e30f, -1e30f);
        for (int t = 0; t < nt; t++) {
            Tri tr;
            V3* vs[3] = {&tr.v0, &tr.v1, &tr.v2};
            for (int k = 0; k < 3; k++) {
                int vi = idx[t * 3 + k];
                if (vi < 0 || vi >= nv) throw std::runtime_error("bad triangle vertex index in collision mesh");
                *vs[k] = V3(vtx[vi * 3 + 0], vtx[vi * 3 + 1], vtx[vi * 3 + 2]);
                meshMn = vmin(meshMn, *vs[k]); meshMx = vmax(meshMx, *vs[k]);
            }
            local.push_back(tr);
        }
        // btQuantizedBvh::setQuantizationValues(meshAabb, margin 1.0) + quantize/unQuantize of the leaf boxes: the
        // reference sorts leaves by the centres of the QUANTISED boxes, which decides ties on regular meshes.
        V3 qMin, qMax, qScale;
        auto quant = [&](V3 p, int isMax, unsigned short* o) {
            V3 v = (p - qMin) * qScale;
            for (int a = 0; a < 3; a++) o[a] = isMax ? (unsigned short)(((unsigned short)(v[a] + 1.f)) | 1) : (unsigned short)(((unsigned short)(v[a])) & 0xfffe);
        };
        auto unquant = [&](const unsigned short* q) { return V3((float)q[0] / qScale.x, (float)q[1] / qScale.y, (float)q[2] / qScale.z) + qMin; };
        {
            V3 clampV(1.f, 1.f, 1.f);
            qMin = meshMn - clampV; qMax = meshMx + clampV;
            V3 sz = qMax - qMin; qScale = V3(65533.f / sz.x, 65533.f / sz.y, 65533.f / sz.z);
            unsigned short q[3];
            quant(qMin, 0, q); qMin = vmin(qMin, unquant(q) - clampV);
            sz = qMax - qMin; qScale = V3(65533.f / sz.x, 65533.f / sz.y, 65533.f / sz.z);
            quant(qMax, 1, q); qMax = vmax(qMax, unquant(q) + clampV);
            sz = qMax - qMin; qScale = V3(65533.f / sz.x, 65533.f / sz.y, 65533.f / sz.z);
        }
        for (int t = 0; t < nt; t++) {
            const Tri& tr = local[t];
            out.tris.push_back(tr);
            detail::Leaf l;
            l.mn = vmin(vmin(tr.v0, tr.v1), tr.v2); l.mx = vmax(vmax(tr.v0, tr.v1), tr.v2);
            const float MIN_DIM = 0.002f, MIN_HALF = 0.001f;  // btOptimizedBvh.cpp NodeTriangleCallback
            for (int a = 0; a < 3; a++) if (l.mx[a] - l.mn[a] < MIN_DIM) { l.mx[a] = l.mx[a] + MIN_HALF; l.mn[a] = l.mn[a] - MIN_HALF; }
            unsigned short q0[3], q1[3];
            quant(l.mn, 0, q0); quant(l.mx, 1, q1);
            l.mn = unquant(q0); l.mx = unquant(q1);
            l.tri = triBase + t;
            bb.leaves.push_back(l);
        }
        ms.nodeStart[m] = (int)out.nodes.size();
        ms.hdrStart[m] = (int)out.hdrRoot.size();
        bb.build(0, nt);
        ms.nodeCount[m] = (int)out.nodes.size() - ms.nodeStart[m];
        if ((int)out.hdrRoot.size() == ms.hdrStart[m]) {  // whole tree fits one header
            out.hdrRoot.push_back(ms.nodeStart[m]);
            out.hdrSize.push_back(ms.nodeCount[m]);
        }
    }
    ms.hdrStart[n] = (int)out.hdrRoot.size();
    ms.numTris = (int)out.tris.size(); ms.numNodes = (int)out.nodes.size(); ms.numHdrs = (int)out.hdrRoot.size();
    {   // free box (rl_mesh.h MeshSet::freeMn/freeMx): largest s with [-sX, sX] x [-sY, sY] x (-inf, inf) clear of every leaf box
        const float PAD = 0.05f;  // > the 0.01 ray-walk padding
        float X = 0.f, Y = 0.f;
        for (const BvhNode& nd : out.nodes) if (nd.tri >= 0) {
            X = std::max(X, std::max(std::fabs(nd.mn[0]), std::fabs(nd.mx[0])));
            Y = std::max(Y, std::max(std::fabs(nd.mn[1]), std::fabs(nd.mx[1])));
        }
        auto touches = [&](float s) {
            for (const BvhNode& nd : out.nodes) if (nd.tri >= 0) {
                if (nd.mx[0] + PAD >= -s * X && nd.mn[0] - PAD <= s * X && nd.mx[1] + PAD >= -s * Y && nd.mn[1] - PAD <= s * Y) return true;
            }
            return false;
        };
        ms.freeMn = V3(1, 1, 1); ms.freeMx = V3(-1, -1, -1);
        if (!out.nodes.empty() && X > 0.f && Y > 0.f && !touches(0.f)) {
            float lo = 0.f, hi = 1.f;
            for (int it = 0; it < 24; it++) { float mid = 0.5f * (lo + hi); if (touches(mid)) hi = mid; else lo = mid; }
            ms.freeMn = V3(-lo * X, -lo * Y, -1e30f); ms.freeMx = V3(lo * X, lo * Y, 1e30f);
        }
    }
    out.triFlags.assign(out.tris.size(), 0);
    out.triEdgeAngles.assign(out.tris.size() * 3, 6.283185307179586232f);
    host_build_edge_info(out);
    host_build_leaf_grid(out);
}

// DiscreteAction table (G/Utils/ActionParsers/DiscreteAction.cpp:3-67)
inline int host_build_action_table(float* t) {
    const float RB[2] = {0, 1}, RF[3] = {-1, 0, 1};
    int n = 0;
    auto push = [&](float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
        float v[8] = {a0, a1, a2, a3, a4, a5, a6, a7};
        for (int k = 0; k < 8; k++) t[n * 8 + k] = v[k];
        n++;
    };
    for (float throttle : RF) for (float steer : RF) for (float boost : RB) for (float handbrake : RB) {
        if (boost == 1 && throttle != 1) continue;
        push(throttle, steer, 0, steer, 0, 0, boost, handbrake);
    }
    for (float pitch : RF) for (float yaw : RF) for (float roll : RF) for (float jump : RB) for (float boost : RB) {
        if (jump == 1 && yaw != 0) continue;
        if (pitch == roll && roll == jump && jump == 0) continue;
        float handbrake = (jump == 1) && (pitch != 0 || yaw != 0 || roll != 0);
        push(boost, yaw, pitch, yaw, roll, jump, boost, handbrake);
    }
    return n;
}

inline void host_build_tables(Tables& tb) {
    int n = host_build_action_table(tb.actions);
    if (n != 90) throw std::runtime_error("action table size != 90");
    // RocketSim pad order: 6 big then 28 small (R/RLConst.h:215-253, Arena.cpp:540-558)
    static const float BIG[6][3] = {{-3584, 0, 73}, {3584, 0, 73}, {-3072, 4096, 73}, {3072, 4096, 73}, {-3072, -4096, 73}, {3072, -4096, 73}};
    static const float SMALL[28][3] = {
        {0, -4240, 70}, {-1792, -4184, 70}, {1792, -4184, 70}, {-940, -3308, 70}, {940, -3308, 70}, {0, -2816, 70}, {-3584, -2484, 70},
        {3584, -2484, 70}, {-1788, -2300, 70}, {1788, -2300, 70}, {-2048, -1036, 70}, {0, -1024, 70}, {2048, -1036, 70}, {-1024, 0, 70},
        {1024, 0, 70}, {-2048, 1036, 70}, {0, 1024, 70}, {2048, 1036, 70}, {-1788, 2300, 70}, {1788, 2300, 70}, {-3584, 2484, 70},
        {3584, 2484, 70}, {0, 2816, 70}, {-940, 3308, 70}, {940, 3308, 70}, {-1792, 4184, 70}, {1792, 4184, 70}, {0, 4240, 70}};
    for (int i = 0; i < 6; i++) for (int k = 0; k < 3; k++) tb.padPos[i * 3 + k] = BIG[i][k];
    for (int i = 0; i < 28; i++) for (int k = 0; k < 3; k++) tb.padPos[(6 + i) * 3 + k] = SMALL[i][k];
    for (int i = 0; i < 34 * 3; i++) tb.padPosBT[i] = tb.padPos[i] * UU2BT;
    {   // BoostPadGrid: pad cell + the clamped 3x3 neighbourhood rule of BoostPadGrid::CheckCollision (BoostPadGrid.cpp:5-25)
        const int CELLS_X = 8, CELLS_Y = 10;
        const int CELL_SIZE_X = (int)(4096.f / (CELLS_X / 2)), CELL_SIZE_Y = (int)(5120.f / (CELLS_Y / 2));
        for (int iy = -1; iy <= CELLS_Y; iy++)
            for (int ix = -1; ix <= CELLS_X; ix++) {
                uint64_t mask = 0;
                int lox = ix - 1 > 0 ? ix - 1 : 0, hix = ix + 1 < CELLS_X - 1 ? ix + 1 : CELLS_X - 1;
                int loy = iy - 1 > 0 ? iy - 1 : 0, hiy = iy + 1 < CELLS_Y - 1 ? iy + 1 : CELLS_Y - 1;
                for (int i = 0; i < 34; i++) {
                    int px = (int)(tb.padPos[i * 3] / CELL_SIZE_X + (CELLS_X / 2));
                    int py = (int)(tb.padPos[i * 3 + 1] / CELL_SIZE_Y + (CELLS_Y / 2));
                    if (px >= lox && px <= hix && py >= loy && py <= hiy) mask |= 1ULL << i;
                }
                int cell = (ix + 1) + (CELLS_X + 2) * (iy + 1);
                tb.padCellMask[cell * 2] = (uint32_t)mask; tb.padCellMask[cell * 2 + 1] = (uint32_t)(mask >> 32);
            }
    }
    // CommonValues::BOOST_LOCATIONS (G/Utils/CommonValues.h:40-75)
    static const float LOC[34][2] = {
        {0, -4240}, {-1792, -4184}, {1792, -4184}, {-3072, -4096}, {3072, -4096}, {-940, -3308}, {940, -3308}, {0, -2816}, {-3584, -2484},
        {3584, -2484}, {-1788, -2300}, {1788, -2300}, {-2048, -1036}, {0, -1024}, {2048, -1036}, {-3584, 0}, {-1024, 0}, {1024, 0},
        {3584, 0}, {-2048, 1036}, {0, 1024}, {2048, 1036}, {-1788, 2300}, {1788, 2300}, {-3584, 2484}, {3584, 2484}, {0, 2816},
        {-940, 3310}, {940, 3308}, {-3072, 4096}, {3072, 4096}, {-1792, 4184}, {1792, 4184}, {0, 4240}};
    for (int i = 0; i < 34; i++) {
        int found = -1;
        for (int j = 0; j < 34; j++) {
            float dx = tb.padPos[j * 3] - LOC[i][0], dy = tb.padPos[j * 3 + 1] - LOC[i][1];
            if (dx * dx + dy * dy < 10) { found = j; break; }
        }
        if (found < 0) throw std::runtime_error("boost pad index map: no matching pad");
        tb.padMap[i] = found;
    }
}

inline void host_mutators_default(rlg_mutators& m) {  // MutatorConfig(GameMode::SOCCAR): RLConst.h values
    memset(&m, 0, sizeof(m));
    m.gravity[2] = C::GRAVITY_Z;
    m.car_mass = C::CAR_MASS; m.car_world_friction = C::CARWORLD_FRICTION; m.car_world_restitution = C::CARWORLD_RESTITUTION;
    m.ball_mass = C::BALL_MASS; m.ball_max_speed = C::BALL_MAX_SPEED; m.ball_drag = C::BALL_DRAG;
    m.ball_world_friction = C::BALL_FRICTION; m.ball_world_restitution = C::BALL_RESTITUTION;
    m.jump_accel = C::JUMP_ACCEL; m.jump_immediate_force = C::JUMP_IMMEDIATE_FORCE;
    m.boost_accel_ground = C::BOOST_ACCEL_GROUND; m.boost_accel_air = C::BOOST_ACCEL_AIR; m.boost_used_per_second = C::BOOST_USED_PER_SECOND;
    m.respawn_delay = C::DEMO_RESPAWN_TIME; m.bump_cooldown_time = C::BUMP_COOLDOWN_TIME;
    m.boost_pad_cooldown_big = C::PAD_COOLDOWN_BIG; m.boost_pad_cooldown_small = C::PAD_COOLDOWN_SMALL;
    m.car_spawn_boost_amount = C::BOOST_SPAWN_AMOUNT; m.ball_hit_extra_force_scale = 1.f; m.bump_force_scale = 1.f;
    m.ball_radius = C::BALL_RADIUS; m.goal_base_threshold_y = C::GOAL_THRESHOLD_Y;
}
inline void host_build_simcfg(const rlg_engine_cfg& c, SimCfg& s) {
    std::memset(&s, 0, sizeof(s));
    if (c.num_arenas <= 0) throw std::runtime_error("num_arenas must be > 0");
    if (c.team_size < 1 || c.team_size > 3) throw std::runtime_error("team_size must be 1..3");
    if (c.tick_skip < 1) throw std::runtime_error("tick_skip must be >= 1");
    s.numArenas = c.num_arenas;
    s.spawnOpponents = c.spawn_opponents != 0;
    s.numCars = c.team_size * (s.spawnOpponents ? 2 : 1);
    s.tickSkip = c.tick_skip;
    s.numActions = RLG_NUM_ACTIONS;
    if (c.car_preset < 0 || c.car_preset >= C::kNumCarPresets) throw std::runtime_error("bad car_preset");
    s.carPreset = c.car_preset;
    s.obsKind = c.obs_kind; s.obsMaxPlayers = c.obs_max_players;
    if (s.obsKind == RLG_OBS_PADDED) {
        // DefaultOBSPadded.cpp:41-45
        if (c.team_size - 1 > c.obs_max_players - 1) throw std::runtime_error("DefaultOBSPadded: Too many teammates for OBS");
        if ((s.spawnOpponents ? c.team_size : 0) > c.obs_max_players) throw std::runtime_error("DefaultOBSPadded: Too many opponents for OBS");
        s.obsSize = 51 + 19 * 2 * c.obs_max_players;
    } else {
        s.obsSize = 51 + 19 * s.numCars;
    }
    if (c.num_reward_terms < 0 || c.num_reward_terms > RLG_MAX_REWARD_TERMS) throw std::runtime_error("bad num_reward_terms");
    s.numRewardTerms = c.num_reward_terms;
    for (int i = 0; i < c.num_reward_terms; i++) {
        s.rewards[i].kind = c.reward_terms[i].kind; s.rewards[i].weight = c.reward_terms[i].weight;
        for (int k = 0; k < 11; k++) s.rewards[i].params[k] = c.reward_terms[i].params[k];
    }
    s.zeroSum = c.zero_sum; s.teamSpirit = c.team_spirit; s.opponentScale = c.opponent_scale;
    s.noTouchMaxSteps = c.no_touch_max_steps; s.goalScoreTerminal = c.goal_score_terminal;
    s.stateSetter = c.state_setter; s.randBallSpeed = c.rand_ball_speed; s.randCarSpeed = c.rand_car_speed; s.carsOnGround = c.cars_on_ground;
    // reference player order = unordered_set<Car*> iteration order; for the small sets used this is descending id
    for (int i = 0; i < s.numCars; i++) s.playerOrder[i] = s.numCars - 1 - i;
    rlg_mutators m;
    host_mutators_default(m);
    if (c.mutators_set) m = c.mutators;
    // car_mass: accepted, and without effect — exactly as in the reference's Gym: Gym::Gym calls Arena::SetMutatorConfig BEFORE it adds the cars
    // (G/Gym.cpp:40-49), SetMutatorConfig only re-masses cars that already exist (Arena.cpp:34-40) and Car::_BulletSetup builds every
    // car with RLConst::CAR_MASS_BT (Car.cpp:206-209), so a Gym's cars weigh 180 whatever MutatorConfig::carMass says.
    if (!(m.car_mass > 0.f)) throw std::runtime_error("mutators: car_mass must be positive");
    if (!(m.ball_mass > 0.f)) throw std::runtime_error("mutators: ball_mass must be positive");
    {   // the reference refuses a dynamic object whose AABB diagonal exceeds a broadphase cell (btRSBroadphase.cpp:229-230, cell size 7.4 Bullet
        // units in soccar: "Object AABB size exceeds maximum cell size"); the ball's AABB is its radius + 0.08 on every side
        const float half = m.ball_radius * UU2BT + 0.08f;
        if (!(m.ball_radius >= 10.f) || 2.f * half * 1.7320508f > 7.4f)
            throw std::runtime_error("mutators: ball_radius must be in [10, 102] uu (the reference's broadphase refuses a larger ball)");
    }
    if (m.demo_mode < 0 || m.demo_mode > 2) throw std::runtime_error("mutators: bad demo_mode");
    if (!(m.ball_drag >= 0.f && m.ball_drag < 1.f)) throw std::runtime_error("mutators: ball_drag must be in [0, 1)");
    Mut& u = s.mut;
    u.gravityBT = V3(m.gravity[0] * UU2BT, m.gravity[1] * UU2BT, m.gravity[2] * UU2BT);
    u.gravityX = m.gravity[0]; u.gravityZ = m.gravity[2];
    u.carWorldFriction = m.car_world_friction; u.carWorldRestitution = m.car_world_restitution;
    u.ballWorldFriction = m.ball_world_friction; u.ballWorldRestitution = m.ball_world_restitution;
    u.ballMaxSpeed = m.ball_max_speed; u.jumpAccel = m.jump_accel; u.jumpImmediateForce = m.jump_immediate_force;
    u.boostAccelGround = m.boost_accel_ground; u.boostAccelAir = m.boost_accel_air; u.boostUsedPerSecond = m.boost_used_per_second;
    u.respawnDelay = m.respawn_delay; u.bumpCooldownTime = m.bump_cooldown_time;
    u.padCooldownBig = m.boost_pad_cooldown_big; u.padCooldownSmall = m.boost_pad_cooldown_small;
    u.carSpawnBoost = m.car_spawn_boost_amount; u.ballHitExtraForceScale = m.ball_hit_extra_force_scale; u.bumpForceScale = m.bump_force_scale;
    u.goalBaseThresholdY = m.goal_base_threshold_y;
    u.ballMass = m.ball_mass; u.ballRadius = m.ball_radius;
    u.unlimitedFlips = m.unlimited_flips != 0; u.unlimitedDoubleJumps = m.unlimited_double_jumps != 0;
    u.demoMode = m.demo_mode; u.enableTeamDemos = m.enable_team_demos != 0;
    s.ballDampFactor = powf(1.f - m.ball_drag, kTickTime);                  // btRigidBody::applyDamping
    s.flipZDampFactor = powf(1 - C::FLIP_Z_DAMP_120, kTickTime / (1 / 120.f));  // Car.cpp:753
}

}  // namespace rl
