// engine.cu — the CUDA collection engine behind the C ABI of include/rlgym_b200.h.
//
// Arena state lives in HBM word-transposed (word w of arena a at state[w * A + a]) so a warp's
// 32 arenas move as coalesced 128-byte lines.  The role kernel (k_roles) keeps a block's arenas
// in shared memory for all ticks of the step (tick_skip physics ticks + gym layer), one warp per
// (group of 32 arenas, body role), and writes them back once: the algorithmic HBM traffic per
// arena-step is read S + write S + actions + obs + rewards + done.  The collision meshes/BVH and the lookup tables are shared, read-only
// and L2-resident.  There is no CPU fallback: every entry point that needs the device
// returns RLG_ERR_CUDA when it is not usable.
#include <cuda_runtime.h>

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "rl_convert.h"
#include "rl_host_build.h"
#include "rl_tick.h"

using namespace rl;

static thread_local std::string g_last_error;
static int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
#define CK(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) return fail(RLG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

constexpr int kHbOffloadDefault = 0;  // see HbJob below: measured 3 % slower with the ball warp taking the hitbox passes (profiles/r02k_ab.txt)
struct rlg_engine {
    rlg_engine_cfg cfgIn;
    SimCfg cfg;
    int device;
    int nwords;
    cudaStream_t stream;
    uint32_t* state = nullptr;
    Tables* tables = nullptr;
    MeshSet ms;  // device pointers
    void* meshMem[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool meshesLoaded = false;
    float* obs = nullptr;
    float* reward = nullptr;
    uint8_t* done = nullptr;
    int32_t* actions = nullptr;  // device staging for step_host
    // pinned host staging
    int32_t* hActions = nullptr;
    float* hObs = nullptr;
    float* hReward = nullptr;
    uint8_t* hDone = nullptr;
    Contact* scratch = nullptr;  // per-arena contact segments (rl_collide.h ContactSink)
    cudaStream_t copyStream = nullptr; cudaEvent_t evFirst = nullptr;  // host-buffer step: D2H of the results overlaps ticks 1..
    int32_t* resetCount = nullptr; int32_t* hResetCount = nullptr; int32_t* hResetIds = nullptr; float* hResetObs = nullptr;
    int32_t* dResetIds = nullptr; float* dResetObs = nullptr;  // device views of the two mapped host buffers
    unsigned char* epa = nullptr;  // [block] full-size penetration-depth workspaces of the role kernel
    void* hbJobs = nullptr;        // [arena][car] HbJob records of the role kernel
    int hbOffload = kHbOffloadDefault;
    float* metrics = nullptr;    // [kMetricWords][A]: stepTotal, stepCount (u32), epTotal, epCount (u32), curEpRew, totalSteps (u32)
    int xwords = 0, stride = 0, scratchSlots = 0;
    int arenasPerBlock = 32, groupsPerBlock = 1;
    size_t rolesSmem = 0;
    int barMode = 0, asyncLoad = 1;
    uint32_t *prof = nullptr, *prof2 = nullptr;  // RLG_PHASE_TIMING builds only
    uint64_t launches = 0;
    // host-plugin path: device staging of rlg_engine_export_gamestates (grown on demand)
    int32_t* xIds = nullptr; rlg_car_state* xCars = nullptr; rlg_ball_state* xBalls = nullptr; rlg_gym_state* xGym = nullptr; rlg_gym_player* xPlayers = nullptr;
    int xCap = 0;
    cudaEvent_t evExport = nullptr;
    // rlg_engine_set_state / reset masks: one device staging buffer, grown on demand (no allocation per call); evStage orders its
    // reuse after the last kernel that read it, whatever stream that ran on
    unsigned char* stage = nullptr; size_t stageCap = 0; cudaEvent_t evStage = nullptr; bool stageBusy = false;
    uint32_t* readyFlags = nullptr; uint32_t readySeq = 0;  // per-block completion flags of the last fused step (rlg_engine_step_ready)
};

// the staging buffer, at least `bytes` long, safe to overwrite from stream s
static int stage_acquire(rlg_engine* e, size_t bytes, cudaStream_t s, unsigned char** out) {
    if (!e->evStage) CK(cudaEventCreateWithFlags(&e->evStage, cudaEventDisableTiming));
    if (bytes > e->stageCap) {
        if (e->stageBusy) CK(cudaEventSynchronize(e->evStage));
        e->stageBusy = false;
        if (e->stage) CK(cudaFree(e->stage));
        e->stage = nullptr; e->stageCap = 0;
        size_t cap = bytes < (1u << 20) ? (1u << 20) : bytes + bytes / 2;
        CK(cudaMalloc(&e->stage, cap));
        e->stageCap = cap;
    }
    if (e->stageBusy) CK(cudaStreamWaitEvent(s, e->evStage, 0));
    *out = e->stage;
    return RLG_OK;
}
static int stage_release(rlg_engine* e, cudaStream_t s) {
    CK(cudaEventRecord(e->evStage, s));
    e->stageBusy = true;
    return RLG_OK;
}

// ---- state movement -----------------------------------------------------------------------------
__device__ __forceinline__ void load_arena(ArenaS& s, const uint32_t* __restrict__ buf, int A, int a, int nwords) {
    uint32_t* w = reinterpret_cast<uint32_t*>(&s);
#pragma unroll 4
    for (int i = 0; i < nwords; i++) w[i] = buf[(size_t)i * A + a];
}
__device__ __forceinline__ void store_arena(const ArenaS& s, uint32_t* __restrict__ buf, int A, int a, int nwords) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&s);
#pragma unroll 4
    for (int i = 0; i < nwords; i++) buf[(size_t)i * A + a] = w[i];
}

// ---- kernels --------------------------------------------------------------------------------------
__global__ void k_init(uint32_t* state, SimCfg cfg, int nwords, uint64_t seed, uint64_t base) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= cfg.numArenas) return;
    ArenaS s;
    arena_init(s, cfg.numCars, seed, base + (uint64_t)a, cfg.mut.carSpawnBoost);
    store_arena(s, state, cfg.numArenas, a, nwords);
}

// Gym::Reset on masked arenas (mask == nullptr: all). useSetter = 0 adopts the current arena state.
__global__ void k_reset(uint32_t* state, SimCfg cfg, int nwords, const Tables* __restrict__ tb, const uint8_t* __restrict__ mask,
                        int useSetter, float* __restrict__ obs) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= cfg.numArenas) return;
    if (mask && !mask[a]) return;
    ArenaS s;
    load_arena(s, state, cfg.numArenas, a, nwords);
    if (useSetter) gym_reset(s, cfg); else episode_reset(s, cfg);
    build_obs(s, cfg, *tb, obs + (size_t)a * cfg.numCars * cfg.obsSize);
    store_arena(s, state, cfg.numArenas, a, nwords);
}

__global__ void k_set_state(uint32_t* state, SimCfg cfg, int nwords, const int32_t* __restrict__ ids, int n,
                            const rlg_car_state* __restrict__ cars, const rlg_ball_state* __restrict__ balls,
                            const rlg_pad_state* __restrict__ pads, const int64_t* __restrict__ ticks) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = ids[i];
    ArenaS s;
    load_arena(s, state, cfg.numArenas, a, nwords);
    if (cars) for (int c = 0; c < cfg.numCars; c++) car_from_pod(s.cars[c], cars[(size_t)i * cfg.numCars + c]);
    if (balls) ball_from_pod(s.ball, balls[i]);
    if (pads) for (int p = 0; p < kNumPads; p++) pad_from_pod(s.pads, p, pads[(size_t)i * kNumPads + p]);
    if (ticks && ticks[i] >= 0) set_i64(s.tickLo, s.tickHi, ticks[i]);
    store_arena(s, state, cfg.numArenas, a, nwords);
}

__global__ void k_get_state(const uint32_t* state, SimCfg cfg, int nwords, const int32_t* __restrict__ ids, int n,
                            rlg_car_state* __restrict__ cars, rlg_ball_state* __restrict__ balls, rlg_pad_state* __restrict__ pads,
                            int64_t* __restrict__ ticks) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = ids[i];
    ArenaS s;
    load_arena(s, state, cfg.numArenas, a, nwords);
    if (cars) for (int c = 0; c < cfg.numCars; c++) {
        rlg_car_state o;
        memset(&o, 0, sizeof(o));
        car_to_pod(o, s.cars[c], c, cfg.spawnOpponents);
        cars[(size_t)i * cfg.numCars + c] = o;
    }
    if (balls) { rlg_ball_state o; ball_to_pod(o, s.ball); balls[i] = o; }
    if (pads) for (int p = 0; p < kNumPads; p++) { rlg_pad_state o; pad_to_pod(o, s.pads, p); pads[(size_t)i * kNumPads + p] = o; }
    if (ticks) ticks[i] = get_i64(s.tickLo, s.tickHi);
}

// GameState::UpdateFromArena's view of n arenas for host plugins (G/Utils/Gamestates/GameState.cpp:52-104, PlayerData.cpp:4-33):
// the CarStates in PLAYER order, the ball, and the gym-layer members (match counters, touch flags, prevActions, score line, pads
// in CommonValues::BOOST_LOCATIONS order).  Run on the state as the step's snapshot left it (tick 0 of a split step, or a reset).
__global__ void k_export_gamestates(const uint32_t* state, SimCfg cfg, int nwords, const Tables* __restrict__ tb, const int32_t* __restrict__ ids, int n,
                                    rlg_car_state* __restrict__ cars, rlg_ball_state* __restrict__ balls, rlg_gym_state* __restrict__ gym,
                                    rlg_gym_player* __restrict__ players) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = ids ? ids[i] : i;
    ArenaS s;
    load_arena(s, state, cfg.numArenas, a, nwords);
    const int64_t tick = get_i64(s.tickLo, s.tickHi);
    for (int p = 0; p < cfg.numCars; p++) {
        const int ci = cfg.playerOrder[p];
        const CarS& c = s.cars[ci];
        if (cars) {
            rlg_car_state o;
            memset(&o, 0, sizeof(o));
            car_to_pod(o, c, ci, cfg.spawnOpponents);
            cars[(size_t)i * cfg.numCars + p] = o;
        }
        if (players) {
            rlg_gym_player o;
            o.match_goals = c.matchGoals; o.match_saves = c.matchSaves; o.match_assists = c.matchAssists; o.match_shots = c.matchShots;
            o.match_shot_passes = c.matchShotPasses; o.match_bumps = c.matchBumps; o.match_demos = c.matchDemos; o.boost_pickups = c.boostPickups;
            o.ball_touched_step = c.touchedStep;
            o.ball_touched_tick = c.hitValid ? (get_i64(c.hitTickLo, c.hitTickHi) == tick - 1) : 0;  // PlayerData.cpp:22
            for (int k = 0; k < 8; k++) o.prev_action[k] = c.prevAction[k];
            players[(size_t)i * cfg.numCars + p] = o;
        }
    }
    if (balls) { rlg_ball_state o; ball_to_pod(o, s.ball); balls[i] = o; }
    if (gym) {
        rlg_gym_state o;
        o.tick_count = tick;
        o.score_line[0] = s.scoreLine[0]; o.score_line[1] = s.scoreLine[1]; o.last_touch_car_id = s.lastTouchCarId;
        o.steps_since_touch = s.stepsSinceTouch;
        for (int g = 0; g < kNumPads; g++) {
            const int pi = tb->padMap[g];
            o.pad_active[g] = (int32_t)((pads_active(s.pads) >> pi) & 1ULL);
            o.pad_cooldown[g] = s.pads.cooldown[pi];
        }
        gym[i] = o;
    }
}

// ---- the role kernel: Arena::Step x n (mode 0) or the fused Gym::Step + GameInst auto-reset (mode 1) -----------------------
// One block = G groups of 32 consecutive arenas x (1 + numCars) warps; a warp is ROLE r (0 = ball / arena bookkeeping,
// 1 + c = car c) of one group's 32 arenas, lane = arena.  The block's arenas live in shared memory for the whole launch (arena words +
// per-tick exchange, per-lane stride odd -> conflict-free), are loaded and stored cooperatively as coalesced 128-byte
// lines of the word-transposed HBM layout, and the phases of rl_tick.h are separated by __syncthreads().
struct RolesArgs {
    uint32_t* state;
    SimCfg cfg;
    int nwords, xwords, stride;  // arena words, exchange words, per-lane shared stride (words, odd)
    int arenasPerBlock;          // arenas of one block = its arena groups x 32 lanes (the last group may be partial)
    MeshSet ms;
    const Tables* tb;
    Contact* scratch;
    int scratchSlots;
    int mode;  // 0: tick, 1: step
    int barMode;    // 0: block-wide barriers, 1: per-group named barriers, 2: per-group inside a tick + one block barrier per tick
    int asyncLoad;  // 1: state words move HBM -> shared memory with cp.async
    uint32_t* prof; // RLG_PHASE_TIMING builds: [block][warp][kProfSlots] cycle counters
    CarConsts k;    // car preset constants and contact thresholds: read from the kernel-parameter constant bank
    Thresholds thr; // instead of a per-thread local-memory copy (the wheel arrays are indexed dynamically)
    const rlg_controls* controls; int nticks;
    // mode 1 may run a sub-range [tickBegin, tickEnd) of the step's ticks: the host-buffer step launches tick 0 (+ the gym
    // layer) and the rest separately so the obs / reward / done copy to the host overlaps the remaining ticks
    int tickBegin, tickEnd;
    // auto-reset bookkeeping of that split step: arenas re-set at the end of the step (their obs row changes after the
    // first launch's copy) are listed here — count in device memory, ids / obs rows in mapped page-locked host memory
    int32_t* resetCount; int32_t* resetIds; float* resetObs;
    const int32_t* actions; float* obs; float* reward; uint8_t* done; int autoReset;
    float* metrics;  // GameInst reward metrics, word-transposed [kMetricWords][A] (nullptr: not tracked)
    unsigned char* epa;  // [block][kEpaFullBytes] full-size penetration-depth workspaces (rl_epa.h), used when a warp's small one overflows
    uint32_t* ready; uint32_t readySeq;  // [block] <- readySeq when the block has stored everything (nullptr: not published)
    const uint32_t* actReady; uint32_t actSeq; int actTileRows;  // wait for the producer's tiles that cover this block's action rows (nullptr: none)
    const uint32_t* prevReady; uint32_t prevSeq;  // wait for THIS block of the launch before (the split host-buffer step: tick 0 | ticks 1..)
    struct HbJob* hbJobs;  // [arena][car] hitbox-narrowphase hand-over records
    int hbOffload;         // cars whose hitbox-mesh narrowphase the ball warp runs (0: every car its own, no extra barrier)
};
constexpr int kMetricWords = 6;

constexpr int kProfSlots = 10;  // load, s0, p1, p2, p3, p4+gym, reset+store, barrier wait, total, unused
#ifdef RLG_PHASE_TIMING
#define PT_DECL() uint32_t pt[kProfSlots]; for (int i = 0; i < kProfSlots; i++) pt[i] = 0; long long ptStart = clock64(), ptT = ptStart
#define PT_WORK(i) do { long long n_ = clock64(); pt[i] += (uint32_t)(n_ - ptT); ptT = n_; } while (0)
#else
#define PT_DECL() do {} while (0)
#define PT_WORK(i) do {} while (0)
#endif

__device__ __forceinline__ void bar_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void cp_async4(uint32_t* smemDst, const uint32_t* gmemSrc) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmemSrc) : "memory");
}

// ---- warp-cooperative mesh passes of a car-role warp --------------------------------------------------------------------
// In a car-role warp only a few lanes (cars near the mesh) have candidate triangles, and scanning a candidate list is a
// chain of dependent loads (list entry -> BVH leaf -> triangle) followed by the longest serial stretches of the tick
// (ray-triangle tests for four wheels, support-plane early out + GJK / SAT for the hitbox).  Instead of every such lane
// walking its own list while the rest of the warp idles, the warp queues its (car, candidate) pairs in shared memory —
// a lane's pairs stay contiguous and in candidate order — and evaluates them 32 at a time, one pair per lane, every lane
// in the same code; each car's own lane then folds its results in candidate order, exactly like the serial scans
// (wheel_mesh_rays, box_meshes_candidates).  Cars whose pairs do not fit the queue, or whose candidate list overflowed,
// take the serial path.  All 32 lanes of the warp must call these (active = the lane has a car to process); cx / cw /
// w are only meaningful on active lanes, k, mine (this lane's arena slot), wq and stride on all of them.
constexpr int kWqItems = 128;                      // pairs queued per warp and pass
constexpr int kWqResWords = 8;
constexpr int kWqWords = kWqItems + 32 * kWqResWords;
// one small penetration-depth workspace per block (rl_epa.h: lock word + storage), behind the queues.  Shared memory is
// budgeted to the byte: at 16 384 arenas a block's slots + queues + this stay under the 196 KiB carve-out step, which leaves
// the L1 its 60 KiB for the per-thread stacks (the next step, 228 KiB, costs 0.19 ms per launch).
constexpr int kEpaSmallWords = (int)(kEpaSmallBytes / 4);

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += u; }
    return v;
}

// Pass 0: the candidate lists themselves — the leaf-grid cell lists of the cars near the mesh are scanned one (car, leaf)
// pair per lane; the owning lane appends the overlapping leaves in list order (== collect_candidates).
__device__ __forceinline__ void collect_candidates_warp(CarW& w, int ci, bool active, const CarConsts& k, const MeshSet& ms, const uint32_t* mine,
                                                        uint32_t* wq) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    V3 mn, mx;
    int first = 0, count = 0;
    bool useGrid = false, walk = false;
    if (active) {
        car_cands_box(reinterpret_cast<const ArenaS*>(mine)->cars[ci], k, mn, mx);
        w.cands.n = 0;
        if (!inside_free_box(ms, mn, mx)) {
            if (grid_lookup(ms, mn, mx, first, count)) useGrid = true; else walk = true;
        }
    }
    const int n = useGrid ? count : 0;
    const int incl = warp_incl_scan(n, lane);
    const int base = incl - n;
    const bool fits = useGrid && incl <= kWqItems;
    if (useGrid && !fits) walk = true;
    const unsigned fitLanes = __ballot_sync(full, fits && n > 0);
    if (fitLanes) {
        const int totalFit = __shfl_sync(full, incl, 31 - __clz(fitLanes));
        float* box = reinterpret_cast<float*>(wq + kWqItems);  // per owner lane: query box (6) + first list index (1)
        uint32_t* ev = wq + kWqItems + 32 * 7;                 // per worker lane: the list entry it tested this round
        if (fits && n > 0) {
            float* b = box + lane * 7;
            b[0] = mn.x; b[1] = mn.y; b[2] = mn.z; b[3] = mx.x; b[4] = mx.y; b[5] = mx.z; b[6] = __int_as_float(first);
            for (int j = 0; j < n; j++) wq[base + j] = (uint32_t)lane | ((uint32_t)j << 5);
        }
        __syncwarp();
        for (int r0 = 0; r0 < totalFit; r0 += 32) {
            const int g = r0 + lane;
            bool pass = false;
            if (g < totalFit) {
                const uint32_t it = wq[g];
                const float* b = box + (it & 31u) * 7;
                const int e = ms.gridList[__float_as_int(b[6]) + (int)(it >> 5)];
                const BvhNode& nd = ms.nodes[e & 0xffffff];
                pass = aabb_overlap(nd.mn, nd.mx, V3(b[0], b[1], b[2]), V3(b[3], b[4], b[5]));
                ev[lane] = (uint32_t)e;
            }
            const unsigned bits = __ballot_sync(full, pass);
            __syncwarp();
            if (fits && n > 0 && base < r0 + 32 && base + n > r0) {
                const int lo = (base > r0 ? base : r0) - r0, hi = (base + n < r0 + 32 ? base + n : r0 + 32) - r0;  // own lanes of this round
                unsigned m = bits & (hi >= 32 ? full : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                for (; m; m &= m - 1) {
                    if (w.cands.n < 0) break;
                    if (w.cands.n >= kMaxCands) { w.cands.n = -1; break; }
                    w.cands.node[w.cands.n++] = (int)ev[__ffs(m) - 1];
                }
            }
            __syncwarp();
        }
    }
    if (walk) collect_candidates(ms, mn, mx, w.cands);
    RL_PT(0);
}

// Pass 1 (right after the candidates are collected): the mesh part of the four wheel rays -> w.meshHit, and the hitbox
// pre-filter (leaf box vs hitbox AABB) -> w.candMask / w.candGroupStart.
__device__ __forceinline__ void cands_pass_warp(CarW& w, int ci, bool active, const CarConsts& k, const MeshSet& ms, const uint32_t* mine, uint32_t* wq,
                                                int stride) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    RL_PT(-1);
    if (active) wheel_mesh_rays_init(w);
    const int n = (active && w.cands.n > 0) ? w.cands.n : 0;
    const int incl = warp_incl_scan(n, lane);
    const int base = incl - n;
    const bool fits = active && w.cands.n >= 0 && incl <= kWqItems;
    const unsigned fitLanes = __ballot_sync(full, fits && n > 0);
    if (fitLanes) {
        const int totalFit = __shfl_sync(full, incl, 31 - __clz(fitLanes));  // the fitting lanes' pairs are a prefix of the queue
        if (fits) {
            uint32_t gs = 0;
            for (int j = 0; j < n; j++) {
                const int e = w.cands.node[j];
                wq[base + j] = (uint32_t)(e & 0xffffff) | ((uint32_t)lane << 24);
                if (j == 0 || (e >> 24) != (w.cands.node[j - 1] >> 24)) gs |= 1u << j;
            }
            w.candGroupStart = gs;
        }
        __syncwarp();
        RL_PT(15);
        uint32_t* res = wq + kWqItems;
        int next = 0;  // own candidates folded so far
        uint32_t mask = 0;
        for (int r0 = 0; r0 < totalFit; r0 += 32) {
            const int g = r0 + lane;
            if (g < totalFit) {
                const uint32_t it = wq[g];
                const int owner = (int)(it >> 24);
                const ArenaS& so = *reinterpret_cast<const ArenaS*>(mine + (owner - lane) * stride);
                const CarS& c = so.cars[ci];
                const BvhNode& nd = ms.nodes[it & 0xffffff];
                V3 boxCenter = c.pos + c.rot * k.hitboxOffset;
                V3 ext(dot(vabs(c.rot.r[0]), k.halfExt), dot(vabs(c.rot.r[1]), k.halfExt), dot(vabs(c.rot.r[2]), k.halfExt));
                uint32_t flags = aabb_overlap(nd.mn, nd.mx, boxCenter - ext, boxCenter + ext) ? 16u : 0u;
                uint32_t* r = res + lane * kWqResWords;
                if (!c.isDemoed) {  // Car::_PreTickUpdate returns before the vehicle update
                    const Tri& t = ms.tris[nd.tri];
                    V3 from[4], to[4];
                    wheel_ray_segments(c, k, from, to);
                    TriRays tr;
                    ray_tri4(from, to, t.v0, t.v1, t.v2, tr);
                    r[0] = __float_as_uint(tr.d[0]); r[1] = __float_as_uint(tr.d[1]); r[2] = __float_as_uint(tr.d[2]); r[3] = __float_as_uint(tr.d[3]);
                    r[4] = __float_as_uint(tr.nn.x); r[5] = __float_as_uint(tr.nn.y); r[6] = __float_as_uint(tr.nn.z);
                    flags |= tr.neg;
                } else {
                    r[0] = r[1] = r[2] = r[3] = __float_as_uint(2.f);
                }
                r[7] = flags;
            }
            __syncwarp();
            RL_PT(16);
            if (fits) {
                while (next < n && base + next < r0 + 32) {
                    const uint32_t* r = res + (base + next - r0) * kWqResWords;
                    const float d[4] = {__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3])};
                    ray_tri4_apply(d, V3(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6])), r[7] & 15u, w.meshHit);
                    mask |= ((r[7] >> 4) & 1u) << next;
                    next++;
                }
            }
            __syncwarp();
            RL_PT(17);
        }
        if (fits) w.candMask = mask;
    }
    if (fits) w.haveMask = 1;
    else if (active) wheel_mesh_rays(reinterpret_cast<const ArenaS*>(mine)->cars[ci], k, ms, w);  // serial; haveMask stays 0
    RL_PT(18);
}

// What the hitbox-mesh narrowphase of one car needs from the candidate passes: published by the car's role in P1a (global memory,
// one record per arena and car), consumed in P1b by whichever warp evaluates the pairs — the group's ball warp for the first
// kHbOffload cars (it would idle through the cars' vehicle update otherwise), the car's own warp for the rest.  The consumer
// leaves the car-world callback's result (CollideCtx::wcHas) and nothing else here.
struct HbJob {
    uint32_t candMask, candGroupStart;
    int32_t haveMask, wcHas;
    V3 wcNormal;
    MeshCands cands;
};


// Pass 2: hitbox vs the pre-filtered candidate triangles -> the car's world contact slots.
template <class HbIn>  // HbJob (the hand-over record, when another warp evaluates the pairs) or the car's own CarW
__device__ __forceinline__ void box_meshes_warp(CollideCtx& cx, ContactSink& cw, const MeshSet& ms, const HbIn& w, int ci, float breaking,
                                                bool active, const CarConsts& k, const uint32_t* mine, uint32_t* wq, int stride, const EpaCtx* ws) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const MeshCands& cands = w.cands;
    bool serial = active && !w.haveMask;
    const uint32_t mask = (active && w.haveMask) ? w.candMask : 0u;
    const uint32_t groupStart = mask ? w.candGroupStart : 0u;
    const int cnt = __popc(mask);
    const int incl = warp_incl_scan(cnt, lane);
    const int base = incl - cnt;
    const bool fits = cnt > 0 && incl <= kWqItems;
    if (cnt > 0 && !fits) serial = true;
    const unsigned fitLanes = __ballot_sync(full, fits);
    if (fitLanes) {
        const int totalFit = __shfl_sync(full, incl, 31 - __clz(fitLanes));
        if (fits) {
            int i = base;
            for (uint32_t m = mask; m; m &= m - 1) wq[i++] = (uint32_t)(cands.node[__ffs(m) - 1] & 0xffffff) | ((uint32_t)lane << 24);
        }
        __syncwarp();
        RL_PT(19);
        Manifold m; m.a = 1 + ci; m.b = -1; m.n = 0; m.breaking = breaking;
        uint32_t rem = fits ? mask : 0u;  // own pairs not yet consumed
        int next = base;                  // queue index of the next own pair
        int lastJ = 0;                    // candidate that last fed the open manifold
        uint32_t* res = wq + kWqItems;
        for (int r0 = 0; r0 < totalFit; r0 += 32) {
            const int g = r0 + lane;
            if (g < totalFit) {
                const uint32_t it = wq[g];
                const int owner = (int)(it >> 24);
                const ArenaS& so = *reinterpret_cast<const ArenaS*>(mine + (owner - lane) * stride);
                const CarS& c = so.cars[ci];
                V3 boxCenter = c.pos + c.rot * k.hitboxOffset;
                V3 ext(dot(vabs(c.rot.r[0]), k.halfExt), dot(vabs(c.rot.r[1]), k.halfExt), dot(vabs(c.rot.r[2]), k.halfExt));
                V3 normal, point; float dist = 0.f;
                const bool hit = box_mesh_item(c, k, boxCenter, boxCenter - ext, boxCenter + ext, ms.tris[ms.nodes[it & 0xffffff].tri], breaking, ws, normal, point, dist);
                uint32_t* r = res + lane * kWqResWords;
                r[0] = hit ? 1u : 0u;
                if (hit) {
                    r[1] = __float_as_uint(normal.x); r[2] = __float_as_uint(normal.y); r[3] = __float_as_uint(normal.z);
                    r[4] = __float_as_uint(point.x); r[5] = __float_as_uint(point.y); r[6] = __float_as_uint(point.z);
                    r[7] = __float_as_uint(dist);
                }
            }
            __syncwarp();
            RL_PT(20);
            while (rem && next < r0 + 32) {
                const int j = __ffs(rem) - 1;
                rem &= rem - 1;
                const uint32_t* r = res + (next - r0) * kWqResWords;
                next++;
                if (r[0]) {
                    // a mesh-group boundary in (lastJ, j]: the open manifold belongs to an earlier mesh -> close it
                    // (one manifold per mesh: btCompoundCollisionAlgorithm -> btConvexConcaveCollisionAlgorithm)
                    if (m.n > 0 && (groupStart & ((2u << j) - 1u) & ~((2u << lastJ) - 1u))) { manifold_flush(cw, m); m.n = 0; }
                    lastJ = j;
                    manifold_add(cx, m, V3(__uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3])),
                                 V3(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6])), __uint_as_float(r[7]), &ms,
                                 ms.nodes[cands.node[j] & 0xffffff].tri);
                }
            }
            __syncwarp();
            RL_PT(21);
        }
        if (m.n > 0) manifold_flush(cw, m);
    }
    RL_PT(22);
    if (serial) {
        if (cands.n >= 0) box_meshes_candidates(cx, cw, ms, cands, ci, breaking);
        else box_meshes(cx, cw, ms, ci, breaking);
    }
}

__global__ void __maxnreg__(168) k_roles(const __grid_constant__ RolesArgs g) {
    extern __shared__ uint32_t smem[];
    // warp = (arena group, role): the groups of a block run the same phase at the same time, so a role's instruction
    // stream is fetched once per block and shared by its groups (the kernel is bound by instruction-cache misses).
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int P = g.cfg.numCars, roles = P + 1, A = g.cfg.numArenas;
    const int group = warp / roles, role = warp - group * roles;
    const int slot = group * 32 + lane;
    const int a = blockIdx.x * g.arenasPerBlock + slot;
    const bool valid = slot < g.arenasPerBlock && a < A;
    uint32_t* mine = smem + (size_t)slot * g.stride;
    ArenaS& s = *reinterpret_cast<ArenaS*>(mine);
    TickX x = make_tickx(mine + (g.stride - g.xwords));
    Contact* scratch = g.scratch + (size_t)(valid ? a : 0) * g.scratchSlots;
    // per car-role warp: the pair queue of the cooperative hitbox-mesh narrowphase, behind the arena slots
    uint32_t* wq = smem + (size_t)g.arenasPerBlock * g.stride + (size_t)(group * P + (role > 0 ? role - 1 : 0)) * kWqWords;
    // the groups of a block never touch each other's arenas: every barrier may be group-local (named barrier 1 + group)
    const int barId = 1 + group, barThreads = 32 * roles;
#define SYNC_GROUP() do { PT_WORK(9); if (g.barMode == 0) __syncthreads(); else bar_named(barId, barThreads); PT_WORK(7); } while (0)
#define SYNC_TICK() do { PT_WORK(9); if (g.barMode == 1) bar_named(barId, barThreads); else __syncthreads(); PT_WORK(7); } while (0)
    // penetration-depth workspaces: the block's small one behind the pair queues, its full-size one in global memory
    uint32_t* epaSmall = smem + (size_t)g.arenasPerBlock * g.stride + (size_t)(blockDim.x >> 5) / roles * P * kWqWords;
    const EpaCtx epaCtx = epa_ctx(epaSmall, g.epa + (size_t)blockIdx.x * kEpaFullBytes);
    if (threadIdx.x == 0) epaSmall[0] = 0;  // lock free (ordered before its first use by the barriers below)
    // a kernel launched as this one's programmatic dependent (the collector's inference) may be scheduled from now on: its blocks
    // take the SMs of the role blocks that finish first and wait for the per-block flags below, not for the whole grid
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    PT_DECL();
    if (g.actReady) {  // launched as the programmatic dependent of the inference: its tiles for this block's rows (and with them the previous
                       // step of this block, which those tiles waited for) must be complete before anything is read
        if (threadIdx.x == 0) {
            const int a0 = blockIdx.x * g.arenasPerBlock, a1 = (a0 + g.arenasPerBlock < A ? a0 + g.arenasPerBlock : A) - 1;
            const int t0 = (a0 * P) / g.actTileRows, t1 = (a1 * P + P - 1) / g.actTileRows;
            for (int i = t0; i <= t1; i++) {
                const volatile uint32_t* f = g.actReady + i;
                for (unsigned spin = 0; (int32_t)(*f - g.actSeq) < 0; spin++) {
                    __nanosleep(200);
                    if (spin > (1u << 24)) __trap();  // seconds: the producer is gone
                }
            }
            __threadfence();
        }
        __syncthreads();
    }
    if (g.prevReady) {  // launched as the programmatic dependent of the launch that ran this block's previous ticks: wait for that block only
        if (threadIdx.x == 0) {
            const volatile uint32_t* f = g.prevReady + blockIdx.x;
            for (unsigned spin = 0; (int32_t)(*f - g.prevSeq) < 0; spin++) {
                __nanosleep(100);
                if (spin > (1u << 25)) __trap();
            }
            __threadfence();
        }
        __syncthreads();
    }
    if (valid) {
        if (g.asyncLoad) {
            for (int w = role; w < g.nwords; w += roles) cp_async4(mine + w, g.state + (size_t)w * A + a);
            asm volatile("cp.async.wait_all;" ::: "memory");
        } else {
            for (int w = role; w < g.nwords; w += roles) mine[w] = g.state[(size_t)w * A + a];
        }
    }
    PT_WORK(0);
    SYNC_TICK();

    const CarConsts& k = g.k;
    const Thresholds& thr = g.thr;
    CarW w;  // this car role's per-tick work state (wheel contacts, forces)
    bool doneFlag = false;

    if (valid) {
        if (g.mode == 0) {
            if (role > 0 && g.controls) s.cars[role - 1].controls = controls_from(g.controls[(size_t)a * P + (role - 1)]);
        } else if (role == 0 && g.tickBegin == 0) {
            int32_t act[kMaxCars];
            for (int p = 0; p < P; p++) {
                int v = g.actions[(size_t)a * P + p];
                act[p] = v < 0 ? 0 : (v >= g.cfg.numActions ? g.cfg.numActions - 1 : v);
            }
            parse_actions(s, g.cfg, *g.tb, act);
        }
    }
    const int tBegin = g.mode == 0 ? 0 : g.tickBegin, tEnd = g.mode == 0 ? g.nticks : g.tickEnd;
    for (int t = tBegin; t < tEnd; t++) {
        const int first = (g.mode == 1 && t == 0) ? 1 : 0;
        if (valid) { if (role == 0) tick_s0_ball(s, x); else tick_s0_car(s, x, role - 1); }
        PT_WORK(1);
        if (g.barMode == 2 && t > tBegin) SYNC_TICK(); else SYNC_GROUP();  // B1
        // P1a: the ball's own narrowphase | the cars' mesh candidates, wheel-ray mesh part and hitbox pre-filter
        HbJob* jobs = g.hbJobs + (size_t)(valid ? a : 0) * P;
        if (role == 0) {
            if (valid) tick_p1_ball(s, x, g.cfg, g.ms, k, thr, scratch);
        } else {
            if (valid) tick_p1_car_pose(s, x, g.cfg, g.ms, k, role - 1, w, false);
            collect_candidates_warp(w, role - 1, valid, k, g.ms, mine, wq);  // whole warp
            cands_pass_warp(w, role - 1, valid, k, g.ms, mine, wq, g.stride);  // whole warp
            if (valid && role - 1 < (P < g.hbOffload ? P : g.hbOffload)) {  // hand the hitbox narrowphase's input over (only to another warp)
                HbJob& j = jobs[role - 1];
                j.candMask = w.candMask; j.candGroupStart = w.candGroupStart; j.haveMask = w.haveMask; j.wcHas = 0;
                j.cands.n = w.cands.n;
                for (int q = 0; q < w.cands.n; q++) j.cands.node[q] = w.cands.node[q];
            }
        }
        PT_WORK(2);
        const int nOff = P < g.hbOffload ? P : g.hbOffload;
        if (nOff > 0) SYNC_GROUP();  // B1b
        // P1b: the cars' vehicle / control model, car-ball, hitbox-plane | the hitbox-mesh pairs of the first cars on the ball warp
        if (role == 0) {
            for (int c = 0; c < nOff; c++) {
                CollideCtx cx; ContactSink cw;
                HbJob& j = jobs[c];
                if (valid) {
                    cx.a = &s; cx.cfg = &g.cfg; cx.tx = x; cx.k = &k; cx.tick = get_i64(s.tickLo, s.tickHi); cx.firstTickOfStep = first; cx.epa = &epaCtx;
                    cx.wcHas = &j.wcHas; cx.wcNormal = &j.wcNormal;
                    cw = make_sink(seg_car_world(scratch, c), kSegCarWorld);
                }
                box_meshes_warp(cx, cw, g.ms, j, c, thr.car, valid, k, mine, wq, g.stride, &epaCtx);  // whole warp (car 0's queue: its warp is not in a pass now)
                if (valid) car_set_mesh_count(x.car[c], cw.n);
            }
        } else {
            CollideCtx cx; ContactSink cw;
            if (valid) tick_p1_car_begin(s, x, g.cfg, g.ms, k, thr, role - 1, w, scratch, first, cx, cw, &epaCtx);
            if (role - 1 >= nOff) {
                box_meshes_warp(cx, cw, g.ms, w, role - 1, thr.car, valid, k, mine, wq, g.stride, &epaCtx);  // whole warp, from the car's own work state
                if (valid) car_set_mesh_count(x.car[role - 1], cw.n);
            }
            if (valid) tick_p1_car_end(cx, x, thr, role - 1, scratch);
        }
        PT_WORK(2);
        SYNC_GROUP();  // B2
        bool self = false;
        if (valid) {
            if (role == 0) tick_p2_solve(s, x, g.cfg, k, thr, scratch, first);
            else {
                if (role - 1 < nOff) car_world_merge(s, x, role - 1, jobs[role - 1].wcHas, jobs[role - 1].wcNormal);
                if ((self = tick_p2_car_self(s, x, g.cfg, k, role - 1, scratch))) tick_p3_car(s, x, *g.tb, k, role - 1, w);
            }
        }
        PT_WORK(3);
        SYNC_GROUP();  // B3
        if (valid && role > 0 && !self) tick_p3_car(s, x, *g.tb, k, role - 1, w);
        PT_WORK(4);
        SYNC_GROUP();  // B4
        if (valid && role == 0) {
            tick_p4_pads(s, x, g.cfg);
            if (first) {  // Gym::Step after its first tick (G/Gym.cpp:84-93)
                float* o = g.obs + (size_t)a * P * g.cfg.obsSize;
                event_tracker_update(s, g.cfg);
                snapshot_update(s, g.cfg);
                build_obs(s, g.cfg, *g.tb, o);
                doneFlag = compute_done(s, g.cfg);
                compute_rewards(s, g.cfg, g.reward + (size_t)a * P);
                g.done[a] = doneFlag ? 1 : 0;
                if (g.metrics) {  // GameInst::Step (GameInst.cpp:13-31)
                    float totalRew = 0;
                    for (int p = 0; p < P; p++) totalRew += g.reward[(size_t)a * P + p];
                    float* m = g.metrics + a;
                    uint32_t* mu = reinterpret_cast<uint32_t*>(m);
                    if (!isnan(totalRew)) { m[0] += totalRew; mu[(size_t)1 * A] += (uint32_t)P; }
                    float cur = m[(size_t)4 * A] + totalRew / P;
                    if (doneFlag) {
                        if (!isnan(cur)) { m[(size_t)2 * A] += cur; mu[(size_t)3 * A] += 1u; }
                        cur = 0;
                    }
                    m[(size_t)4 * A] = cur;
                    mu[(size_t)5 * A] += 1u;
                }
            }
        }
        PT_WORK(5);
    }
    if (valid && role == 0 && g.mode == 1 && g.autoReset && tEnd == g.cfg.tickSkip) {  // GameInst::Step auto-reset (GameInst.cpp:20-24)
        if (tBegin > 0) doneFlag = g.done[a] != 0;  // decided by the launch that ran tick 0
        if (doneFlag) {
            float* o = g.obs + (size_t)a * P * g.cfg.obsSize;
            gym_reset(s, g.cfg);
            build_obs(s, g.cfg, *g.tb, o);
            if (g.resetCount) {
                const int row = P * g.cfg.obsSize;
                const int i = atomicAdd(g.resetCount, 1);
                g.resetIds[i] = a;
                float* dst = g.resetObs + (size_t)i * row;
                for (int q = 0; q < row; q++) dst[q] = o[q];
            }
        }
    }
    PT_WORK(6);
    SYNC_GROUP();
    if (valid)
        for (int w2 = role; w2 < g.nwords; w2 += roles) g.state[(size_t)w2 * A + a] = mine[w2];
    if (g.ready) {  // everything this block writes is stored: publish (writers -> barrier -> one cumulative device-scope fence -> flag)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(g.ready + blockIdx.x) = g.readySeq;
        }
    }
    PT_WORK(6);
#ifdef RLG_PHASE_TIMING
    if (g.prof && lane == 0) {
        pt[8] = (uint32_t)(clock64() - ptStart);
        uint32_t* dst = g.prof + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * kProfSlots;
        for (int i = 0; i < kProfSlots; i++) dst[i] += pt[i];
    }
#endif
#undef SYNC_GROUP
#undef SYNC_TICK
}

// Match::BuildObservations / IsDone / GetRewards on the CURRENT arena state (Gym::Step minus the physics and the
// event tracker): used to prove the gym layer bit-exact on states injected from the reference.
__global__ void k_eval(uint32_t* state, SimCfg cfg, int nwords, const Tables* __restrict__ tb, const int32_t* __restrict__ actions,
                       float* __restrict__ obs, float* __restrict__ reward, uint8_t* __restrict__ done) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= cfg.numArenas) return;
    ArenaS s;
    load_arena(s, state, cfg.numArenas, a, nwords);
    int32_t act[kMaxCars];
    for (int p = 0; p < cfg.numCars; p++) {
        int v = actions[(size_t)a * cfg.numCars + p];
        act[p] = v < 0 ? 0 : (v >= cfg.numActions ? cfg.numActions - 1 : v);
    }
    parse_actions(s, cfg, *tb, act);
    snapshot_update(s, cfg);
    build_obs(s, cfg, *tb, obs + (size_t)a * cfg.numCars * cfg.obsSize);
    bool d = compute_done(s, cfg);
    compute_rewards(s, cfg, reward + (size_t)a * cfg.numCars);
    done[a] = d ? 1 : 0;
    store_arena(s, state, cfg.numArenas, a, nwords);
}

// ---- host side -----------------------------------------------------------------------------------
static inline int grid_for(int n, int block) { return (n + block - 1) / block; }

extern "C" {

const char* rlg_last_error(void) { return g_last_error.c_str(); }
void rlg_internal_set_error(const char* msg) { g_last_error = msg ? msg : ""; }  // used by collector.cu
int rlg_abi_version(void) { return 1; }
size_t rlg_sizeof_car_state(void) { return sizeof(rlg_car_state); }
size_t rlg_sizeof_engine_cfg(void) { return sizeof(rlg_engine_cfg); }

void rlg_engine_cfg_default(rlg_engine_cfg* c) {
    memset(c, 0, sizeof(*c));
    c->num_arenas = 256; c->team_size = 1; c->spawn_opponents = 1; c->tick_skip = 8; c->device = 0; c->seed = 123;
    c->obs_kind = RLG_OBS_DEFAULT; c->obs_max_players = 3;
    c->num_reward_terms = 4;
    c->reward_terms[0].kind = RLG_REW_FACE_BALL; c->reward_terms[0].weight = 0.1f;
    c->reward_terms[1].kind = RLG_REW_VEL_PLAYER_TO_BALL; c->reward_terms[1].weight = 0.5f;
    c->reward_terms[2].kind = RLG_REW_VEL_BALL_TO_GOAL; c->reward_terms[2].weight = 1.0f;
    c->reward_terms[3].kind = RLG_REW_EVENT; c->reward_terms[3].weight = 50.f;
    c->reward_terms[3].params[1] = 1.f; c->reward_terms[3].params[2] = -1.f;
    c->opponent_scale = 1.f;
    c->no_touch_max_steps = 150; c->goal_score_terminal = 1;
    c->state_setter = RLG_SETTER_RANDOM; c->rand_ball_speed = c->rand_car_speed = c->cars_on_ground = 1;
    host_mutators_default(c->mutators); c->mutators_set = 0;
}

void rlg_mutators_default(rlg_mutators* m) { if (m) host_mutators_default(*m); }

int rlg_action_table(float* table_host) {
    if (!table_host) return fail(RLG_ERR_INVALID, "null table");
    host_build_action_table(table_host);
    return RLG_OK;
}

int rlg_engine_destroy(rlg_engine* e) {
    if (!e) return RLG_OK;
    cudaSetDevice(e->device);
#ifdef RLG_EPA_TIMING
    {   // diagnostic build: time spent in the penetration-depth search
        unsigned long long t[8] = {0};
        cudaMemcpyFromSymbol(t, g_epa_timing, sizeof(t));
        fprintf(stderr, "[epa timing] calls %llu cycles %llu (%.0f per call), role launches %d\n", t[1], t[0], t[1] ? (double)t[0] / (double)t[1] : 0.0, (int)e->launches);
        fprintf(stderr, "[epa timing] guesses %llu; GJK: %llu evaluations, %llu iterations, %llu cycles; EPA: %llu iterations, %llu cycles\n", t[7], t[3], t[4], t[2], t[6], t[5]);
    }
#endif
#ifdef RLG_PHASE_TIMING
    if (e->prof) {  // diagnostic build: dump the per-warp phase cycle counters
        const int blocks = (e->cfg.numArenas + e->arenasPerBlock - 1) / e->arenasPerBlock, warps = e->groupsPerBlock * (1 + e->cfg.numCars);
        std::vector<uint32_t> h((size_t)blocks * warps * kProfSlots);
        cudaMemcpy(h.data(), e->prof, h.size() * 4, cudaMemcpyDeviceToHost);
        const char* path = getenv("RLG_PHASE_DUMP");
        if (FILE* f = fopen(path ? path : "phase_prof.bin", "wb")) {
            int32_t hdr[4] = {blocks, warps, kProfSlots, 1 + e->cfg.numCars};
            fwrite(hdr, 4, 4, f); fwrite(h.data(), 4, h.size(), f);
            std::vector<uint32_t> h2((size_t)blocks * warps * 32);
            cudaMemcpy(h2.data(), e->prof2, h2.size() * 4, cudaMemcpyDeviceToHost);
            fwrite(h2.data(), 4, h2.size(), f);
            fclose(f);
        }
        cudaFree(e->prof); cudaFree(e->prof2);
    }
#endif
    cudaFree(e->scratch); cudaFree(e->metrics); cudaFree(e->epa); cudaFree(e->hbJobs); cudaFree(e->stage); cudaFree(e->readyFlags);
    if (e->evStage) cudaEventDestroy(e->evStage);
    cudaFree(e->xIds); cudaFree(e->xCars); cudaFree(e->xBalls); cudaFree(e->xGym); cudaFree(e->xPlayers);
    if (e->evExport) cudaEventDestroy(e->evExport);
    cudaFree(e->state); cudaFree(e->tables); cudaFree(e->obs); cudaFree(e->reward); cudaFree(e->done); cudaFree(e->actions);
    for (void* p : e->meshMem) cudaFree(p);
    cudaFreeHost(e->hActions); cudaFreeHost(e->hObs); cudaFreeHost(e->hReward); cudaFreeHost(e->hDone);
    cudaFreeHost(e->hResetCount); cudaFreeHost(e->hResetIds); cudaFreeHost(e->hResetObs); cudaFree(e->resetCount);
    if (e->copyStream) cudaStreamDestroy(e->copyStream);
    if (e->evFirst) cudaEventDestroy(e->evFirst);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return RLG_OK;
}

int rlg_engine_create(const rlg_engine_cfg* cfg, rlg_engine** out) {
    if (!cfg || !out) return fail(RLG_ERR_INVALID, "null argument");
    *out = nullptr;
    rlg_engine* e = new (std::nothrow) rlg_engine();
    if (!e) return fail(RLG_ERR_INVALID, "out of host memory");
    try {
        host_build_simcfg(*cfg, e->cfg);
    } catch (std::exception& ex) {
        delete e;
        return fail(RLG_ERR_INVALID, ex.what());
    }
    e->cfgIn = *cfg;
    e->device = cfg->device;
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev <= 0) { delete e; return fail(RLG_ERR_CUDA, "no CUDA device available: the engine has no CPU fallback"); }
    if (cfg->device < 0 || cfg->device >= ndev) { delete e; return fail(RLG_ERR_INVALID, "bad device ordinal"); }
#define CKD(expr)                                                                                       \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) { std::string m = std::string(#expr) + ": " + cudaGetErrorString(_e); rlg_engine_destroy(e); return fail(RLG_ERR_CUDA, m); } \
    } while (0)
    CKD(cudaSetDevice(e->device));
    CKD(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    const int A = e->cfg.numArenas, P = e->cfg.numCars;
    e->nwords = arena_words(P);
    e->xwords = tickx_words(P);
    e->stride = e->nwords + e->xwords;
    if ((e->stride & 1) == 0) e->stride++;  // odd per-lane stride: the 32 lanes of a warp hit 32 different banks
    e->scratchSlots = contact_scratch_slots(P);
    {   // block shape of the role kernel: one block per SM holding ceil(A / SMs) arenas (all SMs busy, one wave), capped
        // by the shared memory one block may use
        cudaDeviceProp prop;
        CKD(cudaGetDeviceProperties(&prop, e->device));
        int perSm = (A + prop.multiProcessorCount - 1) / prop.multiProcessorCount;
        // per arena slot: its words + its share of the car warps' pair queues; slots are allocated in whole groups of 32
        // (the last group may be partial: keep one group's queues in reserve)
        int maxBySmem = (int)((prop.sharedMemPerBlockOptin - 1024 - (size_t)P * kWqWords * 4 - kEpaSmallBytes) / (((size_t)e->stride + (size_t)P * kWqWords / 32) * 4));
        cudaFuncAttributes fa;
        CKD(cudaFuncGetAttributes(&fa, k_roles));
        int maxThreads = prop.regsPerBlock / (fa.numRegs > 0 ? fa.numRegs : 1);
        if (maxThreads > 1024) maxThreads = 1024;
        int maxByThreads = (maxThreads / (32 * (1 + P))) * 32;
        int cap = maxBySmem < maxByThreads ? maxBySmem : maxByThreads;
        if (cap < 1) cap = 1;
        int waves = (perSm + cap - 1) / cap;              // blocks each SM runs one after the other
        int apb = (perSm + waves - 1) / waves;            // ... of equal size
        if (apb < 32) {  // small pools: one group per block with 8 or 16 of a warp's lanes in use spreads the arenas over more SMs and shortens
                         // the per-warp chain (fewer divergent paths per warp): 2 048 arenas 0.75 -> 0.65 ms, 1 024 arenas 0.70 -> 0.57 ms
                         // (profiles/r02zh_small_pools_ab.txt); sizes that are not a power of two were slower
            int p2 = 8;
            while (p2 < apb) p2 <<= 1;
            apb = p2 < cap ? p2 : cap;
        }
        if (const char* ev = getenv("RLG_ARENAS_PER_BLOCK")) apb = atoi(ev);  // profiling A/B only
        if (apb > cap) apb = cap;
        if (apb < 1) apb = 1;
        e->arenasPerBlock = apb;
        e->groupsPerBlock = (apb + 31) / 32;
    }
    if (const char* ev = getenv("RLG_HB_OFFLOAD")) e->hbOffload = atoi(ev);  // profiling A/B only
    if (const char* ev = getenv("RLG_BARRIER_MODE")) e->barMode = atoi(ev);
    if (const char* ev = getenv("RLG_ASYNC_LOAD")) e->asyncLoad = atoi(ev);
    if (e->groupsPerBlock > 15) e->barMode = 0;  // 16 hardware barriers per block
#ifdef RLG_PHASE_TIMING
    {
        const int blocks = (A + e->arenasPerBlock - 1) / e->arenasPerBlock, warps = e->groupsPerBlock * (1 + P);
        CKD(cudaMalloc(&e->prof, (size_t)blocks * warps * kProfSlots * 4));
        CKD(cudaMemset(e->prof, 0, (size_t)blocks * warps * kProfSlots * 4));
        CKD(cudaMalloc(&e->prof2, (size_t)blocks * warps * 32 * 4));
        CKD(cudaMemset(e->prof2, 0, (size_t)blocks * warps * 32 * 4));
        CKD(cudaMemcpyToSymbol(g_rl_pt, &e->prof2, sizeof(e->prof2)));
    }
#endif
    e->rolesSmem = ((size_t)e->arenasPerBlock * e->stride + (size_t)e->groupsPerBlock * P * kWqWords) * 4 + kEpaSmallBytes;
    {   // the attribute belongs to the FUNCTION (per device), not to this engine: a later, smaller engine (the SkillTracker's eval pool) must not
        // lower it under an earlier engine's blocks — keep the high-water mark of every engine created on the device
        static size_t s_rolesSmemMax[64] = {0};
        const int d = e->device >= 0 && e->device < 64 ? e->device : 0;
        if (e->rolesSmem > s_rolesSmemMax[d]) {
            CKD(cudaFuncSetAttribute(k_roles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->rolesSmem));
            s_rolesSmemMax[d] = e->rolesSmem;
        }
    }
    if (const char* cv = getenv("RLG_SMEM_CARVEOUT")) CKD(cudaFuncSetAttribute(k_roles, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv)));
    CKD(cudaMalloc(&e->state, (size_t)e->nwords * A * 4));
    CKD(cudaMalloc(&e->scratch, (size_t)A * e->scratchSlots * sizeof(Contact)));
    {
        const size_t blocks = (A + e->arenasPerBlock - 1) / e->arenasPerBlock;
        CKD(cudaMalloc(&e->readyFlags, blocks * 4));
        CKD(cudaMemsetAsync(e->readyFlags, 0, blocks * 4, e->stream));
    }
    {   // one full-size penetration-depth workspace per block of the role kernel's grid (their locks start free)
        const size_t blocks = (A + e->arenasPerBlock - 1) / e->arenasPerBlock;
        CKD(cudaMalloc(&e->epa, blocks * kEpaFullBytes));
        CKD(cudaMemsetAsync(e->epa, 0, blocks * kEpaFullBytes, e->stream));
    }
    CKD(cudaMalloc(&e->hbJobs, (size_t)A * P * sizeof(HbJob)));
    CKD(cudaMemsetAsync(e->hbJobs, 0, (size_t)A * P * sizeof(HbJob), e->stream));
    CKD(cudaMalloc(&e->metrics, (size_t)kMetricWords * A * 4));
    CKD(cudaMemsetAsync(e->metrics, 0, (size_t)kMetricWords * A * 4, e->stream));
    CKD(cudaMalloc(&e->tables, sizeof(Tables)));
    CKD(cudaMalloc(&e->obs, (size_t)A * P * e->cfg.obsSize * 4));
    CKD(cudaMalloc(&e->reward, (size_t)A * P * 4));
    CKD(cudaMalloc(&e->done, (size_t)A));
    CKD(cudaMalloc(&e->actions, (size_t)A * P * 4));
    CKD(cudaMallocHost(&e->hActions, (size_t)A * P * 4));
    CKD(cudaMallocHost(&e->hObs, (size_t)A * P * e->cfg.obsSize * 4));
    CKD(cudaMallocHost(&e->hReward, (size_t)A * P * 4));
    CKD(cudaMallocHost(&e->hDone, (size_t)A));
    CKD(cudaStreamCreateWithFlags(&e->copyStream, cudaStreamNonBlocking));
    CKD(cudaEventCreateWithFlags(&e->evFirst, cudaEventDisableTiming));
    CKD(cudaMalloc(&e->resetCount, 4));
    CKD(cudaMallocHost(&e->hResetCount, 4));
    CKD(cudaHostAlloc(&e->hResetIds, (size_t)A * 4, cudaHostAllocMapped));
    CKD(cudaHostAlloc(&e->hResetObs, (size_t)A * P * e->cfg.obsSize * 4, cudaHostAllocMapped));
    CKD(cudaHostGetDevicePointer((void**)&e->dResetIds, e->hResetIds, 0));
    CKD(cudaHostGetDevicePointer((void**)&e->dResetObs, e->hResetObs, 0));
    Tables tb;
    try { host_build_tables(tb); } catch (std::exception& ex) { rlg_engine_destroy(e); return fail(RLG_ERR_INVALID, ex.what()); }
    CKD(cudaMemcpyAsync(e->tables, &tb, sizeof(tb), cudaMemcpyHostToDevice, e->stream));
    CKD(cudaMemsetAsync(e->obs, 0, (size_t)A * P * e->cfg.obsSize * 4, e->stream));
    CKD(cudaMemsetAsync(e->reward, 0, (size_t)A * P * 4, e->stream));
    CKD(cudaMemsetAsync(e->done, 0, (size_t)A, e->stream));
    k_init<<<grid_for(A, 128), 128, 0, e->stream>>>(e->state, e->cfg, e->nwords, cfg->seed, (uint64_t)cfg->arena_id_base);
    e->launches++;
    CKD(cudaGetLastError());
    CKD(cudaStreamSynchronize(e->stream));
    memset(&e->ms, 0, sizeof(e->ms));
    *out = e;
    return RLG_OK;
}

int rlg_engine_load_meshes(rlg_engine* e, const void* const* blobs, const size_t* sizes, int n) {
    if (!e || (n > 0 && (!blobs || !sizes))) return fail(RLG_ERR_INVALID, "null argument");
    HostMeshSet hm;
    try { host_build_meshes(blobs, sizes, n, hm); } catch (std::exception& ex) { return fail(RLG_ERR_INVALID, ex.what()); }
    CK(cudaSetDevice(e->device));
    for (void*& p : e->meshMem) { cudaFree(p); p = nullptr; }
    MeshSet ms = hm.meta;
    auto up = [&](int slot, const void* src, size_t bytes, const void** dst) -> cudaError_t {
        if (bytes == 0) { *dst = nullptr; return cudaSuccess; }
        cudaError_t r = cudaMalloc(&e->meshMem[slot], bytes);
        if (r != cudaSuccess) return r;
        *dst = e->meshMem[slot];
        return cudaMemcpyAsync(e->meshMem[slot], src, bytes, cudaMemcpyHostToDevice, e->stream);
    };
    CK(up(0, hm.nodes.data(), hm.nodes.size() * sizeof(BvhNode), (const void**)&ms.nodes));
    CK(up(1, hm.tris.data(), hm.tris.size() * sizeof(Tri), (const void**)&ms.tris));
    CK(up(2, hm.hdrRoot.data(), hm.hdrRoot.size() * 4, (const void**)&ms.hdrRoot));
    CK(up(3, hm.hdrSize.data(), hm.hdrSize.size() * 4, (const void**)&ms.hdrSize));
    CK(up(4, hm.triFlags.data(), hm.triFlags.size() * 4, (const void**)&ms.triFlags));
    CK(up(5, hm.triEdgeAngles.data(), hm.triEdgeAngles.size() * 4, (const void**)&ms.triEdgeAngles));
    CK(up(6, hm.gridRange.data(), hm.gridRange.size() * 4, (const void**)&ms.gridRange));
    CK(up(7, hm.gridList.data(), hm.gridList.size() * 4, (const void**)&ms.gridList));
    CK(cudaStreamSynchronize(e->stream));
    e->ms = ms;
    e->meshesLoaded = true;
    return RLG_OK;
}

static RolesArgs roles_args(rlg_engine* e) {
    RolesArgs g;
    memset(&g, 0, sizeof(g));
    g.state = e->state; g.cfg = e->cfg; g.nwords = e->nwords; g.xwords = e->xwords; g.stride = e->stride;
    g.ms = e->ms; g.tb = e->tables; g.scratch = e->scratch; g.scratchSlots = e->scratchSlots;
    g.arenasPerBlock = e->arenasPerBlock;
    g.barMode = e->barMode; g.asyncLoad = e->asyncLoad; g.prof = e->prof;
    g.k = car_consts(e->cfg.carPreset); g.thr = contact_thresholds(g.k, e->cfg.mut.ballRadius);
    g.epa = e->epa;
    g.hbJobs = reinterpret_cast<HbJob*>(e->hbJobs);
    g.hbOffload = e->hbOffload;
    return g;
}

static cudaStream_t pick(rlg_engine* e, void* stream) { return stream ? (cudaStream_t)stream : e->stream; }

static int do_reset(rlg_engine* e, const uint8_t* mask_host, void* stream, int useSetter, float* obs_out = nullptr) {
    if (!e) return fail(RLG_ERR_INVALID, "null engine");
    if (!e->meshesLoaded) return fail(RLG_ERR_STATE, "rlg_engine_load_meshes must be called first (RocketSim::Init)");
    if (useSetter && e->cfg.stateSetter == RLG_SETTER_HOST) return fail(RLG_ERR_STATE, "host state setter: use rlg_engine_set_state + rlg_engine_reset_current");
    CK(cudaSetDevice(e->device));
    cudaStream_t s = pick(e, stream);
    uint8_t* dmask = nullptr;
    if (mask_host) {
        if (int rc = stage_acquire(e, (size_t)e->cfg.numArenas, s, &dmask)) return rc;
        CK(cudaMemcpyAsync(dmask, mask_host, e->cfg.numArenas, cudaMemcpyHostToDevice, s));
    }
    k_reset<<<grid_for(e->cfg.numArenas, 64), 64, 0, s>>>(e->state, e->cfg, e->nwords, e->tables, dmask, useSetter, obs_out ? obs_out : e->obs);
    e->launches++;
    CK(cudaGetLastError());
    if (dmask) return stage_release(e, s);
    return RLG_OK;
}
int rlg_engine_reset(rlg_engine* e, const uint8_t* mask_host, void* stream) { return do_reset(e, mask_host, stream, 1); }
int rlg_engine_reset_current(rlg_engine* e, const uint8_t* mask_host, void* stream) { return do_reset(e, mask_host, stream, 0); }
int rlg_engine_reset_current_to(rlg_engine* e, const uint8_t* mask_host, float* obs_out, void* stream) { return do_reset(e, mask_host, stream, 0, obs_out); }

int rlg_engine_set_player_order(rlg_engine* e, const int32_t* car_ids_host) {
    if (!e || !car_ids_host) return fail(RLG_ERR_INVALID, "null argument");
    int seen = 0;
    for (int i = 0; i < e->cfg.numCars; i++) {
        int id = car_ids_host[i];
        if (id < 1 || id > e->cfg.numCars || (seen & (1 << id))) return fail(RLG_ERR_INVALID, "player order must be a permutation of car ids 1..P");
        seen |= 1 << id;
    }
    for (int i = 0; i < e->cfg.numCars; i++) e->cfg.playerOrder[i] = car_ids_host[i] - 1;
    return RLG_OK;
}
int rlg_engine_player_order(const rlg_engine* e, int32_t* car_ids_host) {
    if (!e || !car_ids_host) return fail(RLG_ERR_INVALID, "null argument");
    for (int i = 0; i < e->cfg.numCars; i++) car_ids_host[i] = e->cfg.playerOrder[i] + 1;
    return RLG_OK;
}

int rlg_engine_set_state(rlg_engine* e, const int32_t* ids, int n, const rlg_car_state* cars, const rlg_ball_state* balls,
                         const rlg_pad_state* pads, const int64_t* ticks) {
    if (!e || !ids || n < 0) return fail(RLG_ERR_INVALID, "bad argument");
    if (n == 0) return RLG_OK;
    for (int i = 0; i < n; i++) if (ids[i] < 0 || ids[i] >= e->cfg.numArenas) return fail(RLG_ERR_INVALID, "arena id out of range");
    CK(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    const int P = e->cfg.numCars;
    int32_t* dIds = nullptr; rlg_car_state* dCars = nullptr; rlg_ball_state* dBalls = nullptr; rlg_pad_state* dPads = nullptr; int64_t* dTicks = nullptr;
    // one staging buffer for the five pieces (16-byte aligned), no allocation per call
    auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t bIds = up16((size_t)n * 4), bCars = cars ? up16((size_t)n * P * sizeof(rlg_car_state)) : 0, bBalls = balls ? up16((size_t)n * sizeof(rlg_ball_state)) : 0,
                 bPads = pads ? up16((size_t)n * kNumPads * sizeof(rlg_pad_state)) : 0, bTicks = ticks ? up16((size_t)n * 8) : 0;
    unsigned char* st = nullptr;
    if (int rc = stage_acquire(e, bIds + bCars + bBalls + bPads + bTicks, s, &st)) return rc;
    dIds = reinterpret_cast<int32_t*>(st); st += bIds;
    CK(cudaMemcpyAsync(dIds, ids, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    if (cars) { dCars = reinterpret_cast<rlg_car_state*>(st); st += bCars; CK(cudaMemcpyAsync(dCars, cars, (size_t)n * P * sizeof(rlg_car_state), cudaMemcpyHostToDevice, s)); }
    if (balls) { dBalls = reinterpret_cast<rlg_ball_state*>(st); st += bBalls; CK(cudaMemcpyAsync(dBalls, balls, (size_t)n * sizeof(rlg_ball_state), cudaMemcpyHostToDevice, s)); }
    if (pads) { dPads = reinterpret_cast<rlg_pad_state*>(st); st += bPads; CK(cudaMemcpyAsync(dPads, pads, (size_t)n * kNumPads * sizeof(rlg_pad_state), cudaMemcpyHostToDevice, s)); }
    if (ticks) { dTicks = reinterpret_cast<int64_t*>(st); st += bTicks; CK(cudaMemcpyAsync(dTicks, ticks, (size_t)n * 8, cudaMemcpyHostToDevice, s)); }
    k_set_state<<<grid_for(n, 64), 64, 0, s>>>(e->state, e->cfg, e->nwords, dIds, n, dCars, dBalls, dPads, dTicks);
    e->launches++;
    CK(cudaGetLastError());
    if (int rc = stage_release(e, s)) return rc;
    CK(cudaStreamSynchronize(s));
    return RLG_OK;
}

int rlg_engine_get_state(rlg_engine* e, const int32_t* ids, int n, rlg_car_state* cars, rlg_ball_state* balls, rlg_pad_state* pads,
                         int64_t* ticks) {
    if (!e || !ids || n < 0) return fail(RLG_ERR_INVALID, "bad argument");
    if (n == 0) return RLG_OK;
    for (int i = 0; i < n; i++) if (ids[i] < 0 || ids[i] >= e->cfg.numArenas) return fail(RLG_ERR_INVALID, "arena id out of range");
    CK(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    const int P = e->cfg.numCars;
    int32_t* dIds = nullptr; rlg_car_state* dCars = nullptr; rlg_ball_state* dBalls = nullptr; rlg_pad_state* dPads = nullptr; int64_t* dTicks = nullptr;
    CK(cudaMallocAsync(&dIds, (size_t)n * 4, s));
    CK(cudaMemcpyAsync(dIds, ids, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    if (cars) CK(cudaMallocAsync(&dCars, (size_t)n * P * sizeof(rlg_car_state), s));
    if (balls) CK(cudaMallocAsync(&dBalls, (size_t)n * sizeof(rlg_ball_state), s));
    if (pads) CK(cudaMallocAsync(&dPads, (size_t)n * kNumPads * sizeof(rlg_pad_state), s));
    if (ticks) CK(cudaMallocAsync(&dTicks, (size_t)n * 8, s));
    k_get_state<<<grid_for(n, 64), 64, 0, s>>>(e->state, e->cfg, e->nwords, dIds, n, dCars, dBalls, dPads, dTicks);
    e->launches++;
    CK(cudaGetLastError());
    if (cars) CK(cudaMemcpyAsync(cars, dCars, (size_t)n * P * sizeof(rlg_car_state), cudaMemcpyDeviceToHost, s));
    if (balls) CK(cudaMemcpyAsync(balls, dBalls, (size_t)n * sizeof(rlg_ball_state), cudaMemcpyDeviceToHost, s));
    if (pads) CK(cudaMemcpyAsync(pads, dPads, (size_t)n * kNumPads * sizeof(rlg_pad_state), cudaMemcpyDeviceToHost, s));
    if (ticks) CK(cudaMemcpyAsync(ticks, dTicks, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaFreeAsync(dIds, s));
    if (dCars) CK(cudaFreeAsync(dCars, s));
    if (dBalls) CK(cudaFreeAsync(dBalls, s));
    if (dPads) CK(cudaFreeAsync(dPads, s));
    if (dTicks) CK(cudaFreeAsync(dTicks, s));
    CK(cudaStreamSynchronize(s));
    return RLG_OK;
}

int rlg_engine_tick(rlg_engine* e, const rlg_controls* controls, int nticks, void* stream) {
    if (!e || nticks < 0) return fail(RLG_ERR_INVALID, "bad argument");
    if (!e->meshesLoaded) return fail(RLG_ERR_STATE, "rlg_engine_load_meshes must be called first (RocketSim::Init)");
    CK(cudaSetDevice(e->device));
    cudaStream_t s = pick(e, stream);
    RolesArgs g = roles_args(e);
    g.mode = 0; g.controls = controls; g.nticks = nticks;
    k_roles<<<grid_for(e->cfg.numArenas, e->arenasPerBlock), 32 * e->groupsPerBlock * (1 + e->cfg.numCars), e->rolesSmem, s>>>(g);
    e->launches++;
    CK(cudaGetLastError());
    return RLG_OK;
}

static int do_step(rlg_engine* e, const int32_t* action_idx, cudaStream_t s, int autoReset, float* obs = nullptr, float* reward = nullptr,
                   uint8_t* done = nullptr, int tickBegin = 0, int tickEnd = -1, bool listResets = false, const uint32_t* actReady = nullptr,
                   uint32_t actSeq = 0, int actTileRows = 0, int chain = 0 /* 1: publish flags for a chained launch, 2: wait for the launch before */) {
    RolesArgs g = roles_args(e);
    g.tickBegin = tickBegin; g.tickEnd = tickEnd < 0 ? e->cfg.tickSkip : tickEnd;
    if (listResets) { g.resetCount = e->resetCount; g.resetIds = e->dResetIds; g.resetObs = e->dResetObs; }
    g.mode = 1; g.actions = action_idx; g.obs = obs ? obs : e->obs; g.reward = reward ? reward : e->reward; g.done = done ? done : e->done;
    g.autoReset = autoReset;
    g.metrics = autoReset ? e->metrics : nullptr;  // GameInst::Step is the auto-resetting step
    if (chain == 2 && e->readyFlags && e->barMode == 0 && e->readySeq > 0) { g.prevReady = e->readyFlags; g.prevSeq = e->readySeq; }
    if (e->readyFlags && e->barMode == 0 && (g.tickEnd == e->cfg.tickSkip || chain == 1)) { g.ready = e->readyFlags; g.readySeq = ++e->readySeq; }
    if ((actReady && actTileRows > 0 && e->barMode == 0) || g.prevReady) {
        if (actReady) { g.actReady = actReady; g.actSeq = actSeq; g.actTileRows = actTileRows; }
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid_for(e->cfg.numArenas, e->arenasPerBlock)); cfg.blockDim = dim3(32 * e->groupsPerBlock * (1 + e->cfg.numCars));
        cfg.dynamicSmemBytes = e->rolesSmem; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, k_roles, g));
        e->launches++;
        return RLG_OK;
    }
    k_roles<<<grid_for(e->cfg.numArenas, e->arenasPerBlock), 32 * e->groupsPerBlock * (1 + e->cfg.numCars), e->rolesSmem, s>>>(g);
    e->launches++;
    CK(cudaGetLastError());
    return RLG_OK;
}

int rlg_engine_step_to_after(rlg_engine* e, const int32_t* action_idx, float* obs_out, float* reward_out, uint8_t* done_out, void* stream,
                             const uint32_t* tile_flags_dev, uint32_t seq, int rows_per_tile) {
    if (!e || !action_idx || !obs_out || !reward_out || !done_out) return fail(RLG_ERR_INVALID, "null argument");
    if (tile_flags_dev && rows_per_tile < 1) return fail(RLG_ERR_INVALID, "rows_per_tile must be positive");
    if (!e->meshesLoaded) return fail(RLG_ERR_STATE, "rlg_engine_load_meshes must be called first (RocketSim::Init)");
    CK(cudaSetDevice(e->device));
    return do_step(e, action_idx, pick(e, stream), 1, obs_out, reward_out, done_out, 0, -1, false, tile_flags_dev, seq, rows_per_tile);
}

int rlg_engine_step(rlg_engine* e, const int32_t* action_idx, void* stream) {
    if (!e || !action_idx) return fail(RLG_ERR_INVALID, "null argument");
    if (!e->meshesLoaded) return fail(RLG_ERR_STATE, "rlg_engine_load_meshes must be called first (RocketSim::Init)");
    CK(cudaSetDevice(e->device));
    return do_step(e, action_idx, pick(e, stream), 1);
}
int rlg_engine_step_to(rlg_engine* e, const int32_t* action_idx, float* obs_out, float* reward_out, uint8_t* done_out, void* stream) {
    if (!e || !action_idx || !obs_out || !reward_out || !done_out) return fail(RLG_ERR_INVALID, "null argument");
    if (!e->meshesLoaded) return fail(RLG_ERR_STATE, "rlg_engine_load_meshes must be called first (RocketSim::Init)");
    CK(cudaSetDevice(e->device));
    return do_step(e, action_idx, pick(e, stream), 1, obs_out, reward_out, done_out);
}
int rlg_engine_set_action_table(rlg_engine* e, const float* table_host, int n_actions) {
    if (!e || !table_host) return fail(RLG_ERR_INVALID, "null argument");
    if (n_actions < 1 || n_actions > RLG_MAX_ACTIONS) return fail(RLG_ERR_INVALID, "n_actions must be in [1, RLG_MAX_ACTIONS]");
    for (int i = 0; i < n_actions * 8; i++) if (!std::isfinite(table_host[i])) return fail(RLG_ERR_INVALID, "action table holds a non-finite value");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(reinterpret_cast<char*>(e->tables) + offsetof(Tables, actions), table_host, (size_t)n_actions * 8 * 4, cudaMemcpyHostToDevice));
    e->cfg.numActions = n_actions;
    return RLG_OK;
}
int rlg_engine_num_actions(const rlg_engine* e) { return e ? e->cfg.numActions : 0; }
int rlg_engine_step_ready(rlg_engine* e, const uint32_t** flags_dev, uint32_t* seq, int* arenas_per_block) {
    if (!e || !flags_dev || !seq || !arenas_per_block) return fail(RLG_ERR_INVALID, "null argument");
    *flags_dev = (e->barMode == 0 && e->readySeq > 0) ? e->readyFlags : nullptr;
    *seq = e->readySeq; *arenas_per_block = e->arenasPerBlock;
    return RLG_OK;
}
int rlg_engine_metrics(rlg_engine* e, rlg_metrics_host* out) {
    if (!e || !out) return fail(RLG_ERR_INVALID, "null argument");
    CK(cudaSetDevice(e->device));
    const size_t A = (size_t)e->cfg.numArenas;
    std::vector<uint32_t> h((size_t)kMetricWords * A);
    CK(cudaMemcpyAsync(h.data(), e->metrics, h.size() * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    auto f = [&](int w, size_t a) { float v; memcpy(&v, &h[(size_t)w * A + a], 4); return v; };
    // ThreadAgentManager::GetMetrics: AvgTracker += AvgTracker over the games, float totals (AvgTracker.h:34-44)
    float stepTotal = 0, epTotal = 0;
    uint64_t stepCount = 0, epCount = 0, steps = 0;
    for (size_t a = 0; a < A; a++) {
        float st = f(0, a), et = f(2, a);
        if (!std::isnan(st)) { stepTotal += st; stepCount += h[1 * A + a]; }
        if (!std::isnan(et)) { epTotal += et; epCount += h[3 * A + a]; }
        steps += h[5 * A + a];
    }
    out->step_reward_total = stepTotal; out->episode_reward_total = epTotal;
    out->step_reward_count = stepCount; out->episode_count = epCount; out->total_steps = steps;
    out->avg_step_reward = stepCount ? stepTotal / stepCount : NAN;
    out->avg_episode_reward = epCount ? epTotal / epCount : NAN;
    return RLG_OK;
}
int rlg_engine_score_lines(rlg_engine* e, int32_t* out_host) {
    if (!e || !out_host) return fail(RLG_ERR_INVALID, "null argument");
    CK(cudaSetDevice(e->device));
    const size_t A = (size_t)e->cfg.numArenas;
    const size_t w0 = offsetof(ArenaS, scoreLine) / 4;  // word-transposed state: one contiguous row of A words per member word
    std::vector<int32_t> h(2 * A);
    CK(cudaMemcpyAsync(h.data(), e->state + w0 * A, 2 * A * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    for (size_t a = 0; a < A; a++) { out_host[2 * a] = h[a]; out_host[2 * a + 1] = h[A + a]; }
    return RLG_OK;
}
int rlg_engine_reset_metrics(rlg_engine* e) {
    if (!e) return fail(RLG_ERR_INVALID, "null argument");
    CK(cudaSetDevice(e->device));
    // GameInst::ResetMetrics clears the two trackers; curEpRew and totalSteps keep running (GameInst.h)
    CK(cudaMemsetAsync(e->metrics, 0, (size_t)4 * e->cfg.numArenas * 4, e->stream));
    return RLG_OK;
}
int rlg_engine_device(const rlg_engine* e) { return e ? e->device : -1; }
int rlg_engine_arena_id_base(const rlg_engine* e) { return e ? e->cfgIn.arena_id_base : 0; }
int rlg_engine_step_noreset(rlg_engine* e, const int32_t* action_idx, void* stream) {
    if (!e || !action_idx) return fail(RLG_ERR_INVALID, "null argument");
    if (!e->meshesLoaded) return fail(RLG_ERR_STATE, "rlg_engine_load_meshes must be called first (RocketSim::Init)");
    CK(cudaSetDevice(e->device));
    return do_step(e, action_idx, pick(e, stream), 0);
}

int rlg_engine_eval_gym(rlg_engine* e, const int32_t* action_idx, void* stream) {
    if (!e || !action_idx) return fail(RLG_ERR_INVALID, "null argument");
    CK(cudaSetDevice(e->device));
    cudaStream_t s = pick(e, stream);
    k_eval<<<grid_for(e->cfg.numArenas, 64), 64, 0, s>>>(e->state, e->cfg, e->nwords, e->tables, action_idx, e->obs, e->reward, e->done);
    e->launches++;
    CK(cudaGetLastError());
    return RLG_OK;
}

int rlg_engine_outputs(rlg_engine* e, float** obs, float** reward, uint8_t** done) {
    if (!e) return fail(RLG_ERR_INVALID, "null engine");
    if (obs) *obs = e->obs;
    if (reward) *reward = e->reward;
    if (done) *done = e->done;
    return RLG_OK;
}
int rlg_engine_obs_size(const rlg_engine* e) { return e ? e->cfg.obsSize : 0; }
int rlg_engine_num_players(const rlg_engine* e) { return e ? e->cfg.numCars : 0; }
int rlg_engine_num_arenas(const rlg_engine* e) { return e ? e->cfg.numArenas : 0; }
size_t rlg_engine_state_bytes_per_arena(const rlg_engine* e) { return e ? (size_t)e->nwords * 4 : 0; }

int rlg_engine_read_outputs(rlg_engine* e, float* obs_host, float* reward_host, uint8_t* done_host) {
    if (!e) return fail(RLG_ERR_INVALID, "null engine");
    CK(cudaSetDevice(e->device));
    const int A = e->cfg.numArenas, P = e->cfg.numCars;
    if (obs_host) CK(cudaMemcpyAsync(obs_host, e->obs, (size_t)A * P * e->cfg.obsSize * 4, cudaMemcpyDeviceToHost, e->stream));
    if (reward_host) CK(cudaMemcpyAsync(reward_host, e->reward, (size_t)A * P * 4, cudaMemcpyDeviceToHost, e->stream));
    if (done_host) CK(cudaMemcpyAsync(done_host, e->done, (size_t)A, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return RLG_OK;
}

// The host-buffer Gym::Step: H2D action indices -> tick 0 + gym layer -> [D2H obs / reward / done on a second stream]
// overlapped with ticks 1.. + auto-reset -> the obs rows of re-set arenas (written by the kernel straight into mapped
// page-locked memory) are patched into the host obs buffer.  Obs, rewards and done flags are final after the first tick
// of a step (G/Gym.cpp:84-93) except for the arenas GameInst::Step re-sets (GameInst.cpp:20-24), so the 11.8 MB result
// copy (cfg2) hides behind 7/8 of the physics instead of following it.
static int step_pinned_impl(rlg_engine* e, int want_obs) {
    const int A = e->cfg.numArenas, P = e->cfg.numCars;
    const size_t row = (size_t)P * e->cfg.obsSize;
    cudaStream_t s = e->stream, c = e->copyStream;
    CK(cudaMemcpyAsync(e->actions, e->hActions, (size_t)A * P * 4, cudaMemcpyHostToDevice, s));
    if (e->cfg.tickSkip < 2) {  // nothing to overlap with
        int rc = do_step(e, e->actions, s, 1);
        if (rc != RLG_OK) return rc;
        if (want_obs) CK(cudaMemcpyAsync(e->hObs, e->obs, (size_t)A * row * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(e->hReward, e->reward, (size_t)A * P * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(e->hDone, e->done, (size_t)A, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        return RLG_OK;
    }
    CK(cudaMemsetAsync(e->resetCount, 0, 4, s));
    static const bool chainSplit = [] { const char* ev = getenv("RLG_SPLIT_CHAIN"); return !(ev && atoi(ev) == 0); }();
    int rc = do_step(e, e->actions, s, 1, nullptr, nullptr, nullptr, 0, 1, false, nullptr, 0, 0, chainSplit ? 1 : 0);
    if (rc != RLG_OK) return rc;
    CK(cudaEventRecord(e->evFirst, s));
    CK(cudaStreamWaitEvent(c, e->evFirst, 0));
    if (want_obs) CK(cudaMemcpyAsync(e->hObs, e->obs, (size_t)A * row * 4, cudaMemcpyDeviceToHost, c));
    CK(cudaMemcpyAsync(e->hReward, e->reward, (size_t)A * P * 4, cudaMemcpyDeviceToHost, c));
    CK(cudaMemcpyAsync(e->hDone, e->done, (size_t)A, cudaMemcpyDeviceToHost, c));
    // the rest of the step as the programmatic dependent of the first launch: every block goes on as soon as ITS first tick is stored
    rc = do_step(e, e->actions, s, 1, nullptr, nullptr, nullptr, 1, e->cfg.tickSkip, want_obs != 0, nullptr, 0, 0, chainSplit ? 2 : 0);
    if (rc != RLG_OK) return rc;
    CK(cudaMemcpyAsync(e->hResetCount, e->resetCount, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(c));
    CK(cudaStreamSynchronize(s));
    if (want_obs) {
        const int n = *e->hResetCount;
        for (int i = 0; i < n; i++) memcpy(e->hObs + (size_t)e->hResetIds[i] * row, e->hResetObs + (size_t)i * row, row * 4);
    }
    return RLG_OK;
}

int rlg_engine_step_host(rlg_engine* e, const int32_t* action_idx_host, float* obs_host, float* reward_host, uint8_t* done_host) {
    if (!e || !action_idx_host) return fail(RLG_ERR_INVALID, "null argument");
    if (!e->meshesLoaded) return fail(RLG_ERR_STATE, "rlg_engine_load_meshes must be called first (RocketSim::Init)");
    CK(cudaSetDevice(e->device));
    const int A = e->cfg.numArenas, P = e->cfg.numCars;
    memcpy(e->hActions, action_idx_host, (size_t)A * P * 4);
    int rc = step_pinned_impl(e, obs_host != nullptr);
    if (rc != RLG_OK) return rc;
    if (obs_host) memcpy(obs_host, e->hObs, (size_t)A * P * e->cfg.obsSize * 4);
    if (reward_host) memcpy(reward_host, e->hReward, (size_t)A * P * 4);
    if (done_host) memcpy(done_host, e->hDone, (size_t)A);
    return RLG_OK;
}

int rlg_engine_host_buffers(rlg_engine* e, int32_t** action_idx, float** obs, float** reward, uint8_t** done) {
    if (!e) return fail(RLG_ERR_INVALID, "null engine");
    if (action_idx) *action_idx = e->hActions;
    if (obs) *obs = e->hObs;
    if (reward) *reward = e->hReward;
    if (done) *done = e->hDone;
    return RLG_OK;
}

int rlg_engine_step_pinned(rlg_engine* e, int want_obs) {
    if (!e) return fail(RLG_ERR_INVALID, "null engine");
    if (!e->meshesLoaded) return fail(RLG_ERR_STATE, "rlg_engine_load_meshes must be called first (RocketSim::Init)");
    CK(cudaSetDevice(e->device));
    return step_pinned_impl(e, want_obs);
}

// ---- host-plugin path (user OBSBuilder / RewardFunction / TerminalCondition / StepCallback on the host) ---------------------------
// Gym::Step split at the point where the reference takes its GameState (G/Gym.cpp:84-93):
//   rlg_engine_step_begin   parse + first tick + event tracker + GameState::UpdateFromArena (+ the fused built-in obs / reward / done)
//   rlg_engine_export_gamestates_async / _wait   that snapshot to the host (kernel on the engine's stream, D2H on its copy stream)
//   rlg_engine_step_end     the remaining tickSkip - 1 ticks (no auto-reset: the host decides `done`)
int rlg_engine_step_begin(rlg_engine* e, const int32_t* action_idx, float* obs_out, float* reward_out, uint8_t* done_out, void* stream) {
    if (!e || !action_idx) return fail(RLG_ERR_INVALID, "null argument");
    if (!e->meshesLoaded) return fail(RLG_ERR_STATE, "rlg_engine_load_meshes must be called first (RocketSim::Init)");
    CK(cudaSetDevice(e->device));
    return do_step(e, action_idx, pick(e, stream), 0, obs_out, reward_out, done_out, 0, 1);
}
int rlg_engine_step_end(rlg_engine* e, const int32_t* action_idx, void* stream) {
    if (!e || !action_idx) return fail(RLG_ERR_INVALID, "null argument");
    CK(cudaSetDevice(e->device));
    if (e->cfg.tickSkip < 2) return RLG_OK;
    return do_step(e, action_idx, pick(e, stream), 0, nullptr, nullptr, nullptr, 1, e->cfg.tickSkip);
}
int rlg_engine_export_gamestates_async(rlg_engine* e, const int32_t* ids, int n, rlg_car_state* cars, rlg_ball_state* balls, rlg_gym_state* gym,
                                       rlg_gym_player* players) {
    if (!e || n < 0) return fail(RLG_ERR_INVALID, "bad argument");
    const int A = e->cfg.numArenas, P = e->cfg.numCars;
    if (!ids) n = A;
    if (n == 0) return RLG_OK;
    if (ids) for (int i = 0; i < n; i++) if (ids[i] < 0 || ids[i] >= A) return fail(RLG_ERR_INVALID, "arena id out of range");
    CK(cudaSetDevice(e->device));
    cudaStream_t s = e->stream, c = e->copyStream;
    if (!e->xCars) {  // staging for every arena, once
        CK(cudaMalloc(&e->xIds, (size_t)A * 4));
        CK(cudaMalloc(&e->xCars, (size_t)A * P * sizeof(rlg_car_state)));
        CK(cudaMalloc(&e->xBalls, (size_t)A * sizeof(rlg_ball_state)));
        CK(cudaMalloc(&e->xGym, (size_t)A * sizeof(rlg_gym_state)));
        CK(cudaMalloc(&e->xPlayers, (size_t)A * P * sizeof(rlg_gym_player)));
        CK(cudaEventCreateWithFlags(&e->evExport, cudaEventDisableTiming));
        e->xCap = A;
    }
    if (n > e->xCap) return fail(RLG_ERR_INVALID, "more arena ids than arenas");
    CK(cudaStreamSynchronize(c));  // the previous export's copies have left the staging buffers
    if (ids) CK(cudaMemcpyAsync(e->xIds, ids, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    k_export_gamestates<<<grid_for(n, 64), 64, 0, s>>>(e->state, e->cfg, e->nwords, e->tables, ids ? e->xIds : nullptr, n, cars ? e->xCars : nullptr,
                                                      balls ? e->xBalls : nullptr, gym ? e->xGym : nullptr, players ? e->xPlayers : nullptr);
    e->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(e->evExport, s));
    CK(cudaStreamWaitEvent(c, e->evExport, 0));
    if (cars) CK(cudaMemcpyAsync(cars, e->xCars, (size_t)n * P * sizeof(rlg_car_state), cudaMemcpyDeviceToHost, c));
    if (balls) CK(cudaMemcpyAsync(balls, e->xBalls, (size_t)n * sizeof(rlg_ball_state), cudaMemcpyDeviceToHost, c));
    if (gym) CK(cudaMemcpyAsync(gym, e->xGym, (size_t)n * sizeof(rlg_gym_state), cudaMemcpyDeviceToHost, c));
    if (players) CK(cudaMemcpyAsync(players, e->xPlayers, (size_t)n * P * sizeof(rlg_gym_player), cudaMemcpyDeviceToHost, c));
    return RLG_OK;
}
int rlg_engine_export_wait(rlg_engine* e) {
    if (!e) return fail(RLG_ERR_INVALID, "null engine");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->copyStream));
    return RLG_OK;
}
int rlg_engine_export_gamestates(rlg_engine* e, const int32_t* ids, int n, rlg_car_state* cars, rlg_ball_state* balls, rlg_gym_state* gym,
                                 rlg_gym_player* players) {
    int rc = rlg_engine_export_gamestates_async(e, ids, n, cars, balls, gym, players);
    return rc != RLG_OK ? rc : rlg_engine_export_wait(e);
}
// Gym::Reset through the engine's own state setter with the obs rows written into a caller DEVICE buffer [A*P, obs]
int rlg_engine_reset_to(rlg_engine* e, const uint8_t* mask_host, float* obs_out, void* stream) { return do_reset(e, mask_host, stream, 1, obs_out); }
// page-locked host memory for the buffers above (plain malloc'ed memory works too, just slower)
void* rlg_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void rlg_host_free(void* p) { if (p) cudaFreeHost(p); }
void* rlg_device_alloc(rlg_engine* e, size_t bytes) {
    if (!e) { fail(RLG_ERR_INVALID, "null engine"); return nullptr; }
    void* p = nullptr;
    cudaError_t err = cudaSetDevice(e->device);
    if (err == cudaSuccess) err = cudaMalloc(&p, bytes ? bytes : 1);
    if (err != cudaSuccess) { fail(RLG_ERR_CUDA, std::string("rlg_device_alloc: ") + cudaGetErrorString(err)); return nullptr; }
    return p;
}
void rlg_device_free(rlg_engine* e, void* p) {
    if (!p) return;
    if (e) cudaSetDevice(e->device);
    cudaFree(p);
}
void rlg_set_last_error(const char* msg) { g_last_error = msg ? msg : ""; }
size_t rlg_sizeof_gym_state(void) { return sizeof(rlg_gym_state); }
size_t rlg_sizeof_gym_player(void) { return sizeof(rlg_gym_player); }

int rlg_engine_copy_to_host(rlg_engine* e, void* dst_host, const void* src_dev, size_t bytes) {
    if (!e || (bytes && (!dst_host || !src_dev))) return fail(RLG_ERR_INVALID, "null argument");
    CK(cudaSetDevice(e->device));
    CK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return RLG_OK;
}
int rlg_engine_copy_to_device(rlg_engine* e, void* dst_dev, const void* src_host, size_t bytes) {
    if (!e || (bytes && (!dst_dev || !src_host))) return fail(RLG_ERR_INVALID, "null argument");
    CK(cudaSetDevice(e->device));
    CK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return RLG_OK;
}

uint64_t rlg_engine_launch_count(const rlg_engine* e) { return e ? e->launches : 0; }
void* rlg_engine_stream(rlg_engine* e) { return e ? (void*)e->stream : nullptr; }
int rlg_engine_sync(rlg_engine* e) {
    if (!e) return fail(RLG_ERR_INVALID, "null engine");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    return RLG_OK;
}

}  // extern "C"
