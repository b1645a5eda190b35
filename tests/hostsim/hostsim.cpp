// hostsim.cpp — TEST TOOL ONLY. Not shipped, not linked into librlgym_b200.so, never a fallback.
//
// Compiles the SAME device headers (rlgymppo_cpp_b200/csrc/rl_*.h, RL_HD functions) with g++ so
// that the CPU-only test suite (`pytest -m "not gpu"`, no GPU in the dev container) can check the
// kernel logic against the reference oracle / golden vectors before the code ever reaches a B200.
// The product path is rlgymppo_cpp_b200/csrc/engine.cu; it fails loudly without a CUDA device.
#define RL_DEBUG_CONTACTS 1
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../rlgymppo_cpp_b200/csrc/rl_convert.h"
#include "../../rlgymppo_cpp_b200/csrc/rl_host_build.h"
#include "../../rlgymppo_cpp_b200/csrc/rl_tick.h"

using namespace rl;

struct HostSim {
    SimCfg cfg;
    Tables tb;
    HostMeshSet hm;
    MeshSet ms;
    std::vector<ArenaS> arenas;
    std::vector<float> obs, reward;
    std::vector<uint8_t> done;
    std::vector<uint32_t> xw;       // one arena's role exchange (TickX)
    std::vector<Contact> scratch;   // one arena's contact segments
};

static void bind_mesh(HostSim* h) {
    h->ms = h->hm.meta;
    h->ms.nodes = h->hm.nodes.data(); h->ms.tris = h->hm.tris.data();
    h->ms.hdrRoot = h->hm.hdrRoot.data(); h->ms.hdrSize = h->hm.hdrSize.data();
    h->ms.triFlags = h->hm.triFlags.data(); h->ms.triEdgeAngles = h->hm.triEdgeAngles.data();
    if (!h->hm.gridRange.empty()) { h->ms.gridRange = h->hm.gridRange.data(); h->ms.gridList = h->hm.gridList.data(); }
}

extern "C" {

void* hs_create(const rlg_engine_cfg* cfg, const void* const* blobs, const size_t* sizes, int n) {
    try {
        HostSim* h = new HostSim();
        host_build_simcfg(*cfg, h->cfg);
        host_build_tables(h->tb);
        host_build_meshes(blobs, sizes, n, h->hm);
        bind_mesh(h);
        h->arenas.resize(cfg->num_arenas);
        h->xw.assign(tickx_words(h->cfg.numCars), 0u);
        h->scratch.resize(contact_scratch_slots(h->cfg.numCars));
        for (int a = 0; a < cfg->num_arenas; a++) arena_init(h->arenas[a], h->cfg.numCars, cfg->seed, (uint64_t)cfg->arena_id_base + a, h->cfg.mut.carSpawnBoost);
        h->obs.resize((size_t)cfg->num_arenas * h->cfg.numCars * h->cfg.obsSize);
        h->reward.resize((size_t)cfg->num_arenas * h->cfg.numCars);
        h->done.resize(cfg->num_arenas);
        return h;
    } catch (std::exception& e) {
        fprintf(stderr, "hs_create: %s\n", e.what());
        return nullptr;
    }
}
void hs_destroy(void* p) { delete (HostSim*)p; }
int hs_obs_size(void* p) { return ((HostSim*)p)->cfg.obsSize; }
int hs_num_cars(void* p) { return ((HostSim*)p)->cfg.numCars; }
void hs_set_player_order(void* p, const int32_t* carIds) {
    HostSim* h = (HostSim*)p;
    for (int i = 0; i < h->cfg.numCars; i++) h->cfg.playerOrder[i] = carIds[i] - 1;
}

void hs_set_state(void* p, int arena, const rlg_car_state* cars, const rlg_ball_state* ball, const rlg_pad_state* pads, int64_t tick) {
    HostSim* h = (HostSim*)p;
    ArenaS& a = h->arenas[arena];
    if (cars) for (int c = 0; c < h->cfg.numCars; c++) car_from_pod(a.cars[c], cars[c]);
    if (ball) ball_from_pod(a.ball, *ball);
    if (pads) for (int i = 0; i < kNumPads; i++) pad_from_pod(a.pads, i, pads[i]);
    if (tick >= 0) set_i64(a.tickLo, a.tickHi, tick);
}
void hs_get_state(void* p, int arena, rlg_car_state* cars, rlg_ball_state* ball, rlg_pad_state* pads, int64_t* tick) {
    HostSim* h = (HostSim*)p;
    ArenaS& a = h->arenas[arena];
    if (cars) for (int c = 0; c < h->cfg.numCars; c++) { memset(&cars[c], 0, sizeof(rlg_car_state)); car_to_pod(cars[c], a.cars[c], c, h->cfg.spawnOpponents); }
    if (ball) ball_to_pod(*ball, a.ball);
    if (pads) for (int i = 0; i < kNumPads; i++) pad_to_pod(pads[i], a.pads, i);
    if (tick) *tick = get_i64(a.tickLo, a.tickHi);
}
void hs_tick(void* p, int arena, const rlg_controls* controls, int nticks) {
    HostSim* h = (HostSim*)p;
    ArenaS& a = h->arenas[arena];
    if (controls) for (int c = 0; c < h->cfg.numCars; c++) a.cars[c].controls = controls_from(controls[c]);
    for (int t = 0; t < nticks; t++) arena_tick(a, h->cfg, h->ms, h->tb, 0, h->xw.data(), h->scratch.data());
}
// Gym::Reset on one arena using whatever state the arena currently holds (no state setter)
void hs_reset_from_current(void* p, int arena, float* obs) {
    HostSim* h = (HostSim*)p;
    ArenaS& a = h->arenas[arena];
    episode_reset(a, h->cfg);
    build_obs(a, h->cfg, h->tb, obs);
}
void hs_reset(void* p, int arena, float* obs) {
    HostSim* h = (HostSim*)p;
    ArenaS& a = h->arenas[arena];
    gym_reset(a, h->cfg);
    build_obs(a, h->cfg, h->tb, obs);
}
// Gym::Step WITHOUT the auto-reset so that obs of terminal steps can be compared too
void hs_step(void* p, int arena, const int32_t* actions, float* obs, float* reward, uint8_t* done) {
    HostSim* h = (HostSim*)p;
    ArenaS& a = h->arenas[arena];
    parse_actions(a, h->cfg, h->tb, actions);
    arena_tick(a, h->cfg, h->ms, h->tb, 1, h->xw.data(), h->scratch.data());
    event_tracker_update(a, h->cfg);
    snapshot_update(a, h->cfg);
    build_obs(a, h->cfg, h->tb, obs);
    bool d = compute_done(a, h->cfg);
    compute_rewards(a, h->cfg, reward);
    *done = d;
    for (int t = 1; t < h->cfg.tickSkip; t++) arena_tick(a, h->cfg, h->ms, h->tb, 0, h->xw.data(), h->scratch.data());
}
void hs_eval_gym(void* p, int arena, const int32_t* actions, float* obs, float* reward, uint8_t* done) {
    HostSim* h = (HostSim*)p;
    ArenaS& a = h->arenas[arena];
    parse_actions(a, h->cfg, h->tb, actions);
    snapshot_update(a, h->cfg);
    build_obs(a, h->cfg, h->tb, obs);
    bool d = compute_done(a, h->cfg);
    compute_rewards(a, h->cfg, reward);
    *done = d;
}
void hs_gym_state(void* p, int arena, int32_t* score2, int32_t* lastTouch, int32_t* counters /*[P*8] player order*/, uint8_t* touched) {
    HostSim* h = (HostSim*)p;
    ArenaS& a = h->arenas[arena];
    score2[0] = a.scoreLine[0]; score2[1] = a.scoreLine[1]; *lastTouch = a.lastTouchCarId;
    for (int i = 0; i < h->cfg.numCars; i++) {
        const CarS& c = a.cars[h->cfg.playerOrder[i]];
        int32_t* o = counters + i * 8;
        o[0] = c.matchGoals; o[1] = c.matchSaves; o[2] = c.matchAssists; o[3] = c.matchShots;
        o[4] = c.matchShotPasses; o[5] = c.matchBumps; o[6] = c.matchDemos; o[7] = c.boostPickups;
        touched[i] = c.touchedStep;
    }
}
int hs_action_table(float* t) { return host_build_action_table(t); }
int hs_mesh_info(void* p, int32_t* out /*numTris,numNodes,numHdrs*/) {
    HostSim* h = (HostSim*)p;
    out[0] = h->ms.numTris; out[1] = h->ms.numNodes; out[2] = h->ms.numHdrs;
    return h->ms.numMeshes;
}
// debug: contacts of the last tick; rows like ref_arena_dump_contacts (impulses not available -> 0)
int hs_dump_contacts(float* out, int maxRows) {
    int n = 0;
    for (int i = 0; i < g_dbg_contacts.n && n < maxRows; i++) {
        const Contact& c = g_dbg_contacts.c[i];
        float* r = out + n * 16;
        r[0] = c.a <= 0 ? (float)c.a : (float)c.a; r[1] = (float)c.b;
        for (int k = 0; k < 3; k++) { r[2 + k] = c.posA[k]; r[5 + k] = c.posB[k]; r[8 + k] = c.normal[k]; }
        r[11] = c.dist; r[12] = 0; r[13] = c.friction; r[14] = c.restitution; r[15] = (float)c.special;
        n++;
    }
    return n;
}
// debug: wheel rays of the last tick: rows {contact point xyz, normal xyz, suspension length, in contact, ground} per wheel
void hs_dump_wheels(int car, float* out) {
    for (int i = 0; i < 4; i++) {
        const WheelW& w = g_dbg_carw[car].w[i];
        float* r = out + 9 * i;
        for (int k = 0; k < 3; k++) { r[k] = w.contactPoint[k]; r[3 + k] = w.contactNormal[k]; }
        r[6] = w.suspLen; r[7] = (float)w.inContact; r[8] = (float)w.ground;
    }
}
// unit probes of the narrowphase (same argument layout as oracle/ref_harness.cpp ref_probe_*): boxT = origin xyz + basis rows,
// halfExt = the btBoxShape constructor argument (half extents WITH margin); out = {normal xyz, point xyz, depth}
static void probe_box(const float* halfExt, const float* boxT, V3& center, M3& rot, V3& core, float& margin) {
    center = V3(boxT[0], boxT[1], boxT[2]);
    rot = M3(V3(boxT[3], boxT[4], boxT[5]), V3(boxT[6], boxT[7], boxT[8]), V3(boxT[9], boxT[10], boxT[11]));
    // btBoxShape(halfExtents): implicit dimensions = halfExtents - 0.04, THEN setSafeMargin may shrink the margin (btBoxShape.cpp:17-26)
    float mn = fminf_(fminf_(halfExt[0], halfExt[1]), halfExt[2]);
    margin = fminf_(0.04f, 0.1f * mn);
    core = V3(halfExt[0] - 0.04f, halfExt[1] - 0.04f, halfExt[2] - 0.04f);
}
int hs_probe_box_triangle(const float* halfExt, const float* boxT, const float* tri, float breaking, float* out) {
    V3 c, core; M3 r; float margin;
    probe_box(halfExt, boxT, c, r, core, margin);
    Tri t; t.v0 = V3(tri[0], tri[1], tri[2]); t.v1 = V3(tri[3], tri[4], tri[5]); t.v2 = V3(tri[6], tri[7], tri[8]);
    V3 n, p; float d;
    if (!box_triangle_contact(c, r, core, margin, t, breaking, nullptr, n, p, d)) return 0;
    for (int k = 0; k < 3; k++) { out[k] = n[k]; out[3 + k] = p[k]; }
    out[6] = d;
    return 1;
}
int hs_probe_box_sphere(const float* halfExt, const float* boxT, const float* center, float radius, float breaking, float* out) {
    V3 c, core; M3 r; float margin;
    probe_box(halfExt, boxT, c, r, core, margin);
    V3 n, p; float d;
    if (!box_sphere_contact(c, r, core, margin, V3(center[0], center[1], center[2]), radius, breaking, nullptr, n, p, d)) return 0;
    for (int k = 0; k < 3; k++) { out[k] = n[k]; out[3 + k] = p[k]; }
    out[6] = d;
    return 1;
}
int hs_probe_box_box(const float* halfA, const float* tA, const float* halfB, const float* tB, float* out, int cap) {
    V3 ca(tA[0], tA[1], tA[2]), cb(tB[0], tB[1], tB[2]);
    M3 ra(V3(tA[3], tA[4], tA[5]), V3(tA[6], tA[7], tA[8]), V3(tA[9], tA[10], tA[11])), rb(V3(tB[3], tB[4], tB[5]), V3(tB[6], tB[7], tB[8]), V3(tB[9], tB[10], tB[11]));
    // btBoxBoxDetector works on getHalfExtentsWithMargin() = (requested - 0.04) + the safe margin
    V3 c0, coreA, coreB; M3 r0; float mA, mB;
    probe_box(halfA, tA, c0, r0, coreA, mA);
    probe_box(halfB, tB, c0, r0, coreB, mB);
    BoxBoxResult r;
    box_box(ca, ra, coreA + V3(mA, mA, mA), cb, rb, coreB + V3(mB, mB, mB), r);
    for (int i = 0; i < r.n && i < cap; i++) { float* o = out + 7 * i; for (int k = 0; k < 3; k++) { o[k] = r.normal[k]; o[3 + k] = r.point[i][k]; } o[6] = r.depth[i]; }
    return r.n;
}
// ns per box_triangle_contact call on one pose (host timing of the penetration path)
double hs_bench_box_triangle(const float* halfExt, const float* boxT, const float* tri, float breaking, int n) {
    V3 c, core; M3 r; float margin;
    probe_box(halfExt, boxT, c, r, core, margin);
    Tri t; t.v0 = V3(tri[0], tri[1], tri[2]); t.v1 = V3(tri[3], tri[4], tri[5]); t.v2 = V3(tri[6], tri[7], tri[8]);
    V3 nn, p; float d; volatile float sink = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; i++) { V3 cc = c + V3(0, 0, 1e-6f * (i & 7)); if (box_triangle_contact(cc, r, core, margin, t, breaking, nullptr, nn, p, d)) sink += d; }
    return std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count() / n;
}
void hs_epa_small_capacity(int verts, int faces) { g_epa_host_small_verts = verts; g_epa_host_small_faces = faces; }
long hs_epa_overflows() { return g_dbg_epa_overflows; }
void hs_pen_kinds(long* out) { out[0] = g_dbg_pen_kind[0]; out[1] = g_dbg_pen_kind[1]; }
void hs_epa_stats(long* out) { out[0] = g_dbg_epa_calls; out[1] = g_dbg_epa_iters; out[2] = g_dbg_epa_maxface; out[3] = g_dbg_epa_maxsv; for (int i = 0; i < 8; i++) out[4 + i] = g_dbg_epa_hist[i]; }
size_t hs_sizeof_arena() { return sizeof(ArenaS); }

// leaf grid vs stackless BVH walk on random query boxes (centres over the whole arena incl. the goal boxes, half extents
// in [halfMin, halfMax] Bullet units): out = {queries the grid answered, candidate leaves found, mismatching queries}
void hs_grid_check(void* p, int n, uint64_t seed, float halfMin, float halfMax, int64_t* out) {
    HostSim* h = (HostSim*)p;
    MeshSet walk = h->ms; walk.gridRange = nullptr; walk.gridList = nullptr;
    MeshSet nofree = h->ms; nofree.freeMn = V3(1, 1, 1); nofree.freeMx = V3(-1, -1, -1);  // exercise the empty cells too
    walk.freeMn = nofree.freeMn; walk.freeMx = nofree.freeMx;
    uint64_t z = seed;
    auto rnd = [&]() { z += 0x9E3779B97F4A7C15ULL; uint64_t v = z; v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ULL; v = (v ^ (v >> 27)) * 0x94D049BB133111EBULL; v ^= v >> 31; return (float)((v >> 40) * (1.0 / 16777216.0)); };
    out[0] = out[1] = out[2] = 0;
    for (int q = 0; q < n; q++) {
        V3 c((rnd() * 2 - 1) * 4200.f * UU2BT, (rnd() * 2 - 1) * 6100.f * UU2BT, (rnd() * 2150.f - 50.f) * UU2BT);
        V3 hext(halfMin + rnd() * (halfMax - halfMin), halfMin + rnd() * (halfMax - halfMin), halfMin + rnd() * (halfMax - halfMin));
        MeshCands a, b;
        collect_candidates(nofree, c - hext, c + hext, a);
        collect_candidates(walk, c - hext, c + hext, b);
        int first, count;
        if (grid_lookup(nofree, c - hext, c + hext, first, count)) out[0]++;
        if (a.n > 0) out[1] += a.n;
        bool same = a.n == b.n;
        for (int i = 0; same && i < a.n; i++) same = a.node[i] == b.node[i];
        if (!same) out[2]++;
    }
}

}  // extern "C"
