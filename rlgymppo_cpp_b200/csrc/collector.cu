// collector.cu — the device-resident ThreadAgent loop: policy/critic MLP inference on the 5th-gen tensor cores
// (tcgen05.mma kind::tf32, FP32 accumulators in TMEM, weights streamed with the bulk-copy engine), action sampling,
// trajectory ring, GAE and the ExperienceBuffer row export.
//
// Reference: P/private/RLGymPPO_CPP/Threading/ThreadAgent.cpp:24-195 (_RunFunc), PPO/DiscretePolicy.cpp:44-62
// (GetAction), PPO/ValueEstimator.cpp:6-27, Threading/ThreadAgentManager.cpp:16-80 (CollectTimesteps),
// Util/TorchFuncs.cpp:5-52 (ComputeGAE), PPO/ExperienceBuffer.cpp:12-70 (SubmitExperience).
//
// MLP kernel design (k_mlp_infer): one CTA = 128 rows (TMEM lanes) of the [N, obs] observation matrix, 4 warps, thread
// t owns row t in every epilogue.  Activations live in shared memory in the UMMA canonical K-major no-swizzle layout,
// split in 32-wide K blocks (128 rows x 128 B each); weights are pre-packed on the host into the same canonical layout
// per (layer, K block) so one K block of a layer is ONE contiguous cp.async.bulk into a 2-slot ring.  One elected
// thread issues copies (one block ahead) and tcgen05.mma (4 x K=8 per block), commits to mbarriers; the epilogue reads
// the accumulator with tcgen05.ld (32 lanes x 32 columns per warp), applies bias + ReLU and writes the next layer's A
// operand back into shared memory.  The final policy layer's epilogue does softmax / clamp / sample / log-prob in
// registers; the critic's final layer (padded to N=16) yields the value.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/rlgym_b200.h"

extern "C" void rlg_internal_set_error(const char* msg);  // engine.cu: the slot rlg_last_error() reads

namespace {

int failc(int code, const std::string& m);

constexpr int kTileM = 128;
constexpr int kMaxWidth = 256;                              // max layer width (UMMA N <= 256)
constexpr int kBlockK = 32;                                 // K elements per block (tf32: 128 B per row)
constexpr int kABlockBytes = kTileM * kBlockK * 4;          // 16 KB
constexpr int kWSlotBytes = kMaxWidth * kBlockK * 4;        // 32 KB
constexpr int kNumWSlots = 3;
constexpr int kSmemA = (kMaxWidth / kBlockK) * kABlockBytes;  // 128 KB
constexpr int kSmemW = kNumWSlots * kWSlotBytes;            // 64 KB
constexpr int kSmemBar = kSmemA + kSmemW;
constexpr int kSmemBias = kSmemBar + 128;                  // this layer's bias (kMaxWidth floats)
constexpr int kSmemTotal = kSmemBias + kMaxWidth * 4;
constexpr int kThreads = 256;                               // two warpgroups: both read the 128 TMEM lanes, each half of the columns
constexpr int kMaxLayers = RLG_MAX_HIDDEN_LAYERS + 1;
constexpr float kActionMinProb = 1e-11f;                    // DiscretePolicy::ACTION_MIN_PROB

struct MlpLayer {
    const float* w;  // packed: [kPad/32][canonical (nPad x 32) block]
    const float* b;  // [nPad]
    int32_t kPad, nPad;
};
struct MlpNet {
    int32_t numLayers;
    int32_t inDim, outDim;
    MlpLayer layer[kMaxLayers];
};
struct InferArgs {
    MlpNet net[2];  // 0 policy, 1 critic
    const float* obs;
    int32_t nRows, obsDim;
    int32_t* action;
    float* logprob;
    float* value;
    uint64_t seed, counter, rowBase;
    int32_t deterministic;
    float temperature;
    int32_t runNet[2];
    // overlap with the step that produces obs (launched as its programmatic dependent): wait for the role blocks that own this tile's rows
    const uint32_t* ready; uint32_t readySeq; int32_t arenasPerBlock, playersPerArena;
    uint32_t* tileDone; uint32_t tileSeq;  // [tile] <- tileSeq when this CTA has stored its rows' actions / log-probs / values (nullptr: not published)
};

// byte offset of element (r, kk) inside one canonical (R x 32) tf32 block: core matrix = 8 rows x 16 B,
// K-adjacent core matrices 128 B apart (LBO), 8-row groups 1024 B apart (SBO)
__host__ __device__ inline uint32_t canon_off(uint32_t r, uint32_t kk) { return (r >> 3) * 1024u + (kk >> 2) * 128u + (r & 7u) * 16u + (kk & 3u) * 4u; }

// ---- PTX wrappers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded spin: a protocol bug traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 24); it++)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
        "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// round-to-nearest (ties away) to TF32 like cvt.rna.tf32.f32, as two integer ops: the cvt runs on a narrow pipe and was
// 25 % of the kernel's stall samples (profiles/r01h_k_mlp_infer.md); inputs are finite (obs, ReLU outputs)
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE (layout_type 0), version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    const uint64_t lbo = 128 >> 4, sbo = 1024 >> 4;
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// instruction descriptor: D=F32 (bits 4-5 = 1), A=B=TF32 (format 2), K-major both, N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct ChunkIter {  // walks (net, layer, kblock) over the nets that run
    int net, layer, kb;
    __device__ bool valid(const InferArgs& a) const { return net < 2; }
    __device__ void skip_to_valid(const InferArgs& a) {
        while (net < 2 && !a.runNet[net]) net++;
    }
    __device__ void next(const InferArgs& a) {
        kb++;
        if (kb * kBlockK >= a.net[net].layer[layer].kPad) {
            kb = 0; layer++;
            if (layer >= a.net[net].numLayers) { layer = 0; net++; skip_to_valid(a); }
        }
    }
};

__global__ void __launch_bounds__(kThreads, 1) k_mlp_infer(const InferArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sW = smem + kSmemA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);  // [0..S) full, [S..2S) free, [2S] mma_done
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(smem + kSmemBar + 64);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int r = t & (kTileM - 1), half = t >> 7;  // accumulator row (TMEM lane) and column half of this thread
    // the fused step launched as this kernel's programmatic dependent may be scheduled once every CTA of this grid is running
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    float* sBias = reinterpret_cast<float*>(smem + kSmemBias);
    const int row0 = blockIdx.x * kTileM;
    const uint32_t barFull = smem_u32(&bars[0]), barFree = smem_u32(&bars[kNumWSlots]), barDone = smem_u32(&bars[2 * kNumWSlots]);

    if (t == 0) {
        for (int i = 0; i < kNumWSlots; i++) { mbar_init(barFull + 8 * i, 1); mbar_init(barFree + 8 * i, 1); }
        mbar_init(barDone, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmemSlot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmemBase = *tmemSlot;
    const uint32_t tmemLane = tmemBase + ((uint32_t)((warp & 3) * 32) << 16);  // a warp reads TMEM lanes 32 * (warp % 4) ..

    // producer / MMA-issuer bookkeeping (thread 0 only)
    ChunkIter loadIt{0, 0, 0};
    loadIt.skip_to_valid(a);
    uint32_t nLoad = 0, nUse = 0, donePhase = 0;

    if (a.ready) {
        // Launched as the programmatic dependent of the fused step that writes obs: this CTA got the SM of a role block that has finished
        // while the step's slowest blocks are still running.  Its rows belong to the arenas of one or two role blocks: wait for their flags.
        // (Letting a CTA take ANY complete tile instead of its own was measured and is no faster: profiles/r02x_collect_overlap_ab.txt.)
        if (t == 0) {
            const int lastRow = (row0 + kTileM < a.nRows ? row0 + kTileM : a.nRows) - 1;
            const int b0 = (row0 / a.playersPerArena) / a.arenasPerBlock, b1 = (lastRow / a.playersPerArena) / a.arenasPerBlock;
            for (int b = b0; b <= b1; b++) {
                const volatile uint32_t* f = a.ready + b;
                for (unsigned spin = 0; (int32_t)(*f - a.readySeq) < 0; spin++) {
                    __nanosleep(200);
                    if (spin > (1u << 24)) __trap();  // seconds: the producer is gone
                }
            }
            __threadfence();
        }
        __syncthreads();
    }
    for (int net = 0; net < 2; net++) {
        if (!a.runNet[net]) continue;
        const MlpNet& N = a.net[net];
        // ---- stage the observation tile as layer 0's A operand (zero padded to kPad) ----
        {
            const int kPad0 = N.layer[0].kPad;
            // 8 independent loads in flight per thread (the loop was latency bound: 26 % of the stall samples)
            const int total = kTileM * kPad0;
            for (int idx0 = t; idx0 < total; idx0 += kThreads * 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    int idx = idx0 + u * kThreads;
                    int rr = idx / kPad0, k = idx - rr * kPad0;
                    int gr = row0 + rr;
                    v[u] = (idx < total && gr < a.nRows && k < a.obsDim) ? __ldcg(a.obs + (size_t)gr * a.obsDim + k) : 0.f;  // L2: the rows may have been written while this kernel was already running
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    int idx = idx0 + u * kThreads;
                    if (idx < total) {
                        int rr = idx / kPad0, k = idx - rr * kPad0;
                        *reinterpret_cast<float*>(sA + (k >> 5) * kABlockBytes + canon_off(rr, k & 31)) = to_tf32(v[u]);
                    }
                }
            }
        }
        fence_proxy_async();
        __syncthreads();

        for (int l = 0; l < N.numLayers; l++) {
            const MlpLayer& L = N.layer[l];
            const int nkb = L.kPad / kBlockK;
            for (int j = t; j < L.nPad; j += kThreads) sBias[j] = __ldg(L.b + j);
            __syncthreads();
            if (warp == 0) {
                if (lane == 0) {
                    const uint32_t idesc = make_idesc(kTileM, L.nPad);
                    for (int kb = 0; kb < nkb; kb++) {
                        // keep the weight stream one block ahead (across layer and net boundaries)
                        while (nLoad <= nUse + 1 && loadIt.valid(a)) {
                            const MlpLayer& LL = a.net[loadIt.net].layer[loadIt.layer];
                            uint32_t slot = nLoad % kNumWSlots, k = nLoad / kNumWSlots;
                            if (k >= 1) mbar_wait(barFree + 8 * slot, (k - 1) & 1);
                            uint32_t bytes = (uint32_t)LL.nPad * kBlockK * 4;
                            mbar_expect_tx(barFull + 8 * slot, bytes);
                            bulk_g2s(smem_u32(sW + slot * kWSlotBytes), LL.w + (size_t)loadIt.kb * LL.nPad * kBlockK, bytes, barFull + 8 * slot);
                            nLoad++;
                            loadIt.next(a);
                        }
                        uint32_t slot = nUse % kNumWSlots, k = nUse / kNumWSlots;
                        mbar_wait(barFull + 8 * slot, k & 1);
                        tc_fence_after();
                        const uint32_t aBase = smem_u32(sA + kb * kABlockBytes), bBase = smem_u32(sW + slot * kWSlotBytes);
#pragma unroll
                        for (int j = 0; j < kBlockK / 8; j++)
                            tc_mma_tf32(tmemBase, make_desc(aBase + j * 256), make_desc(bBase + j * 256), idesc, (kb > 0 || j > 0) ? 1u : 0u);
                        tc_commit(barFree + 8 * slot);
                        nUse++;
                    }
                    tc_commit(barDone);
                }
                __syncwarp();
            }
            mbar_wait(barDone, donePhase);
            donePhase ^= 1;
            tc_fence_after();

            const bool last = (l == N.numLayers - 1);
            if (!last) {
                // bias + ReLU -> next layer's A operand (K block c of the next layer = columns [32c, 32c+32)); the two
                // warpgroups take alternate column chunks, two chunks in flight per thread
                const int nChunks = L.nPad / 32;
                for (int c = half; c < nChunks; c += 4) {
                    const bool two = c + 2 < nChunks;
                    uint32_t v0[32], v1[32];
                    tc_ld32(tmemLane + c * 32, v0);
                    if (two) tc_ld32(tmemLane + (c + 2) * 32, v1);
                    tc_wait_ld();
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        if (h == 1 && !two) break;
                        const int cc = c + 2 * h;
                        const uint32_t* v = h ? v1 : v0;
                        uint8_t* dst = sA + cc * kABlockBytes + (r >> 3) * 1024 + (r & 7) * 16;
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            const float4 bq = *reinterpret_cast<const float4*>(sBias + cc * 32 + 4 * q);
                            float4 o;
                            o.x = to_tf32(fmaxf(__uint_as_float(v[4 * q + 0]) + bq.x, 0.f));
                            o.y = to_tf32(fmaxf(__uint_as_float(v[4 * q + 1]) + bq.y, 0.f));
                            o.z = to_tf32(fmaxf(__uint_as_float(v[4 * q + 2]) + bq.z, 0.f));
                            o.w = to_tf32(fmaxf(__uint_as_float(v[4 * q + 3]) + bq.w, 0.f));
                            *reinterpret_cast<float4*>(dst + q * 128) = o;
                        }
                    }
                }
            } else if (half != 0) {
                // the heads are one row per thread: the first warpgroup does them
            } else if (net == 1) {
                uint32_t v[16];
                tc_ld16(tmemLane, v);
                tc_wait_ld();
                int gr = row0 + r;
                if (gr < a.nRows && a.value) a.value[gr] = __uint_as_float(v[0]) + sBias[0];
            } else {
                // policy head: softmax(logits / temperature) -> clamp -> sample -> log-prob  (DiscretePolicy.cpp:37-62)
                constexpr int kMaxAct = 96;
                float p[kMaxAct];
                const int nAct = N.outDim;
#pragma unroll
                for (int c = 0; c < kMaxAct / 32; c++) {
                    uint32_t v[32];
                    if (c * 32 < L.nPad) {
                        tc_ld32(tmemLane + c * 32, v);
                        tc_wait_ld();
                    }
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        int j = c * 32 + i;
                        p[j] = (j < nAct) ? (__uint_as_float(v[i]) + sBias[j]) / a.temperature : -INFINITY;
                    }
                }
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < kMaxAct; j++) mx = fmaxf(mx, p[j]);
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < kMaxAct; j++) { p[j] = (j < nAct) ? expf(p[j] - mx) : 0.f; sum += p[j]; }
                float total = 0.f;
#pragma unroll
                for (int j = 0; j < kMaxAct; j++) {
                    if (j < nAct) { p[j] = fminf(fmaxf(p[j] / sum, kActionMinProb), 1.f); total += p[j]; }
                }
                int gr = row0 + r;
                int act = 0;
                float pa = p[0];
                if (a.deterministic) {
#pragma unroll
                    for (int j = 1; j < kMaxAct; j++) if (j < nAct && p[j] > pa) { pa = p[j]; act = j; }
                } else {
                    uint64_t h = splitmix64(a.seed ^ splitmix64(a.counter * 0x9E3779B97F4A7C15ull + (a.rowBase + (uint64_t)gr)));
                    float u = (float)(h >> 40) * (1.0f / 16777216.0f) * total;  // torch.multinomial normalises by the sum
                    float cum = 0.f;
                    bool found = false;
#pragma unroll
                    for (int j = 0; j < kMaxAct; j++) {
                        if (j < nAct) {
                            cum += p[j];
                            if (!found && cum > u) { found = true; act = j; pa = p[j]; }
                            if (!found) { act = j; pa = p[j]; }  // numerical tail: last action
                        }
                    }
                }
                if (gr < a.nRows) {
                    if (a.action) a.action[gr] = act;
                    if (a.logprob) a.logprob[gr] = a.deterministic ? 0.f : logf(pa);
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncthreads();
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(256) : "memory");
    }
    if (a.tileDone && t == 0) {  // (the last epilogue ended with a block barrier: every row's outputs are stored)
        __threadfence();
        *reinterpret_cast<volatile uint32_t*>(a.tileDone + blockIdx.x) = a.tileSeq;
    }
}

// ---- GAE: TorchFuncs::ComputeGAE over the reference's concatenation order --------------------------------------------
// One thread per row n (trajectory of T steps).  values is [T+1][N]; the value "after" a row's last step is the next
// row's first value (reference seam quirk, TorchFuncs.cpp:10) and values[T][N-1] for the last row.
__global__ void k_gae(int T, int N, int P, const float* __restrict__ reward, const uint8_t* __restrict__ done, const float* __restrict__ value,
                      float* __restrict__ adv, float* __restrict__ target, float* __restrict__ ret, float gamma, float lambda, float returnStd,
                      float clipRange) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int A = N / P;
    const int ar = n / P;
    float returnScale = 1 / returnStd;
    if (isnan(returnScale)) returnScale = 0;
    float lastGae = 0.f, lastReturn = 0.f;
    for (int t = T - 1; t >= 0; t--) {
        float d = done[(size_t)t * A + ar] ? 1.f : 0.f;
        float truncated = (t == T - 1) ? (d == 0.f ? 1.f : 0.f) : 0.f;  // ThreadAgentManager.cpp:53
        float fd = 1 - d, ft = 1 - truncated;
        float r = reward[(size_t)t * N + n];
        float normRew;
        if (returnStd != 0) {
            normRew = r * returnScale;
            if (clipRange > 0) normRew = fminf(fmaxf(normRew, -clipRange), clipRange);
        } else {
            normRew = r;
        }
        float v = value[(size_t)t * N + n];
        float nextV = (t == T - 1) ? ((n + 1 < N) ? value[n + 1] : value[(size_t)T * N + n]) : value[(size_t)(t + 1) * N + n];
        float predRet = normRew + gamma * nextV * fd;
        float delta = predRet - v;
        float rr = r + lastReturn * gamma * fd * ft;
        ret[(size_t)t * N + n] = rr;
        lastReturn = rr;
        lastGae = delta + gamma * lambda * fd * ft * lastGae;
        adv[(size_t)t * N + n] = lastGae;
        target[(size_t)t * N + n] = v + lastGae;
    }
}

// ---- ExperienceBuffer row export: T-major ring -> reference order (row i = n*T + t) ----------------------------------
__global__ void k_export_rows(int T, int N, int P, int obsSize, const float* __restrict__ obs, const int32_t* __restrict__ action,
                              const float* __restrict__ logprob, const float* __restrict__ reward, const uint8_t* __restrict__ done,
                              const float* __restrict__ target, const float* __restrict__ adv, float* states, int64_t* actions, float* logProbs,
                              float* rewards, float* nextStates, float* dones, float* truncateds, float* valueTargets, float* advantages) {
    // one warp per (t, n) row; lanes stride over the obs floats
    int warpId = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warpId >= T * N) return;
    int t = warpId / N, n = warpId - t * N;
    size_t i = (size_t)n * T + t;
    const float* src = obs + ((size_t)t * N + n) * obsSize;
    const float* nxt = obs + ((size_t)(t + 1) * N + n) * obsSize;
    if (states) for (int k = lane; k < obsSize; k += 32) states[i * obsSize + k] = src[k];
    if (nextStates) for (int k = lane; k < obsSize; k += 32) nextStates[i * obsSize + k] = nxt[k];
    if (lane == 0) {
        int A = N / P;
        float d = done[(size_t)t * A + n / P] ? 1.f : 0.f;
        if (actions) actions[i] = action[(size_t)t * N + n];
        if (logProbs) logProbs[i] = logprob[(size_t)t * N + n];
        if (rewards) rewards[i] = reward[(size_t)t * N + n];
        if (dones) dones[i] = d;
        if (truncateds) truncateds[i] = (t == T - 1 && d == 0.f) ? 1.f : 0.f;
        if (valueTargets) valueTargets[i] = target[(size_t)t * N + n];
        if (advantages) advantages[i] = adv[(size_t)t * N + n];
    }
}

// mean |returns|, |advantages|, |value targets| of a collect (block sums -> 3 double atomics) and its first nFirst returns in the
// reference's concatenation order (row i = n * T + t)
__global__ void __launch_bounds__(256) k_return_stats(int T, int N, const float* __restrict__ ret, const float* __restrict__ adv, const float* __restrict__ tgt,
                                                      double* __restrict__ sums, float* __restrict__ first, int nFirst) {
    __shared__ float red[3][8];
    const size_t total = (size_t)T * N;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) { s0 += fabsf(ret[i]); s1 += fabsf(adv[i]); s2 += fabsf(tgt[i]); }
    for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; red[2][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.f;
        for (int w = 0; w < 8; w++) t += red[threadIdx.x][w];
        atomicAdd(sums + threadIdx.x, (double)t);
    }
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < nFirst; i += 256) { const int n = i / T, t = i - n * T; first[i] = ret[(size_t)t * N + n]; }
}

// nn.Linear weight [out, in] (leading dimension ldw) -> the inference kernel's packed layout: [kPad / 32] canonical (nPad x 32)
// blocks, TF32-rounded, zero padded; bias -> [nPad]
__global__ void k_pack_layer(const float* __restrict__ W, int ldw, const float* __restrict__ b, int out, int in, int kPad, int nPad, float* __restrict__ dW,
                             float* __restrict__ dB) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kPad * nPad) return;
    const int n = i / kPad, k = i - n * kPad;
    const float v = (n < out && k < in) ? to_tf32(W[(size_t)n * ldw + k]) : 0.f;
    dW[(size_t)(k / kBlockK) * nPad * kBlockK + canon_off(n, k % kBlockK) / 4] = v;
    if (k == 0) dB[n] = n < out ? b[n] : 0.f;
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------------------------
struct rlg_collector {
    rlg_engine* e = nullptr;
    rlg_collector_cfg cfg;
    int device = 0, A = 0, P = 0, N = 0, obs = 0, maxT = 0, T = 0;
    uint64_t rowBase = 0;
    // networks
    struct HostLayer { int in = 0, out = 0, kPad = 0, nPad = 0; float* dW = nullptr; float* dB = nullptr; bool set = false; };
    HostLayer L[2][kMaxLayers];
    int numLayers = 0;
    // ring
    float* dObs = nullptr; int32_t* dAction = nullptr; float* dLogprob = nullptr; float* dReward = nullptr; uint8_t* dDone = nullptr;
    float* dValue = nullptr; float* dAdv = nullptr; float* dTarget = nullptr; float* dRet = nullptr;
    bool haveObs0 = false;
    double* dStats = nullptr;  // rlg_collector_return_stats workspace
    uint64_t stepCounter = 0, launches = 0;
    rlg_reset_hook resetHook = nullptr;
    void* resetUser = nullptr;
    rlg_step_hook stepHook = nullptr;
    void* stepUser = nullptr;
    std::vector<uint8_t> hDone;
    std::vector<int32_t> hIds;
    // optional per-kernel timing of the last collect (CUDA events on the launching stream)
    bool timing = false;
    uint32_t* dTileDone = nullptr; uint32_t tileSeq = 0;  // per-tile completion flags of the collect loop's inferences
    bool chain = false;   // the step after an inference as ITS programmatic dependent too (per-block chains), see rlg_collector_collect
    bool overlap = true;  // inference after a fused step starts as the step's programmatic dependent (RLG_COLLECT_OVERLAP=0: A/B only)
    std::vector<cudaEvent_t> evStep, evInfer;  // pairs (start, end)
    int nStepEv = 0, nInferEv = 0;
};

namespace {
int failc(int code, const std::string& m) { rlg_internal_set_error(m.c_str()); return code; }
#define CKC(expr)                                                                                             \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) return failc(RLG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

inline float round_tf32_host(float x) {  // round-to-nearest (ties away), like cvt.rna.tf32.f32
    uint32_t u; memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return x;
    u += 0x1000u; u &= 0xFFFFE000u;
    float r; memcpy(&r, &u, 4); return r;
}

int fill_net(const rlg_collector* c, int net, MlpNet& out) {
    out.numLayers = c->numLayers;
    out.inDim = c->L[net][0].in;
    out.outDim = c->L[net][c->numLayers - 1].out;
    for (int l = 0; l < c->numLayers; l++) {
        const auto& h = c->L[net][l];
        if (!h.set) return failc(RLG_ERR_STATE, "rlg_collector_set_layer has not been called for every layer");
        out.layer[l].w = h.dW; out.layer[l].b = h.dB; out.layer[l].kPad = h.kPad; out.layer[l].nPad = h.nPad;
    }
    return RLG_OK;
}

int launch_infer(rlg_collector* c, const float* obs, int nRows, uint64_t counter, int32_t* action, float* logprob, float* value, cudaStream_t s,
                 bool afterStep = false, bool publishTiles = false) {
    InferArgs a;
    memset(&a, 0, sizeof(a));
    a.runNet[0] = (action || logprob) ? 1 : 0;
    a.runNet[1] = value ? 1 : 0;
    if (!a.runNet[0] && !a.runNet[1]) return RLG_OK;
    for (int n = 0; n < 2; n++) { int rc = fill_net(c, n, a.net[n]); if (rc != RLG_OK) return rc; }
    a.obs = obs; a.nRows = nRows; a.obsDim = c->obs;
    a.action = action; a.logprob = logprob; a.value = value;
    a.seed = c->cfg.seed; a.counter = counter; a.rowBase = c->rowBase;
    a.deterministic = c->cfg.deterministic; a.temperature = c->cfg.temperature;
    int grid = (nRows + kTileM - 1) / kTileM;
    if (publishTiles && c->overlap && !c->timing && nRows == (int)c->N) {
        if (!c->dTileDone) {
            CKC(cudaMalloc(&c->dTileDone, (size_t)grid * 4));
            CKC(cudaMemsetAsync(c->dTileDone, 0, (size_t)grid * 4, s));
        }
        a.tileDone = c->dTileDone; a.tileSeq = ++c->tileSeq;
    }
    if (afterStep && c->overlap && !c->timing && nRows == (int)c->N) {
        // obs is the output of the fused step launched just before on this stream: start as its programmatic dependent and wait per tile
        // for the role blocks that own the tile's rows, so that inference runs under the tail of the step (its slowest blocks)
        const uint32_t* flags = nullptr; uint32_t seq = 0; int apb = 0;
        if (rlg_engine_step_ready(c->e, &flags, &seq, &apb) != RLG_OK) return failc(RLG_ERR_STATE, rlg_last_error());
        if (flags) {
            a.ready = flags; a.readySeq = seq; a.arenasPerBlock = apb; a.playersPerArena = c->P;
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemTotal; cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            const cudaError_t le = cudaLaunchKernelEx(&cfg, k_mlp_infer, a);
            if (le == cudaSuccess) { c->launches++; return RLG_OK; }
            // a driver without programmatic dependent launch: say so once and go on with ordinary launches (same results, no overlap)
            (void)cudaGetLastError();
            fprintf(stderr, "rlgym_b200: programmatic dependent launch unavailable (%s): the collect loop runs without the inference overlap\n", cudaGetErrorString(le));
            c->overlap = false;
            a.ready = nullptr;
        }
    }
    k_mlp_infer<<<grid, kThreads, kSmemTotal, s>>>(a);
    c->launches++;
    CKC(cudaGetLastError());
    return RLG_OK;
}
}  // namespace

extern "C" {

int rlg_collector_destroy(rlg_collector* c) {
    if (!c) return RLG_OK;
    cudaSetDevice(c->device);
    for (int n = 0; n < 2; n++) for (int l = 0; l < kMaxLayers; l++) { cudaFree(c->L[n][l].dW); cudaFree(c->L[n][l].dB); }
    cudaFree(c->dObs); cudaFree(c->dAction); cudaFree(c->dLogprob); cudaFree(c->dReward); cudaFree(c->dDone);
    cudaFree(c->dValue); cudaFree(c->dAdv); cudaFree(c->dTarget); cudaFree(c->dRet); cudaFree(c->dStats); cudaFree(c->dTileDone);
    for (auto ev : c->evStep) cudaEventDestroy(ev);
    for (auto ev : c->evInfer) cudaEventDestroy(ev);
    delete c;
    return RLG_OK;
}

int rlg_collector_create(rlg_engine* e, const rlg_collector_cfg* cfg, rlg_collector** out) {
    if (!e || !cfg || !out) return failc(RLG_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->num_hidden < 1 || cfg->num_hidden > RLG_MAX_HIDDEN_LAYERS) return failc(RLG_ERR_INVALID, "num_hidden must be 1..4");
    if (cfg->max_steps < 1) return failc(RLG_ERR_INVALID, "max_steps must be >= 1");
    if (!(cfg->temperature > 0.f)) return failc(RLG_ERR_INVALID, "temperature must be > 0");
    for (int i = 0; i < cfg->num_hidden; i++)
        for (int h : {cfg->policy_hidden[i], cfg->critic_hidden[i]})
            if (h < 32 || h > kMaxWidth || (h % 32) != 0) return failc(RLG_ERR_INVALID, "hidden layer sizes must be multiples of 32 in [32, 256]");
    rlg_collector* c = new (std::nothrow) rlg_collector();
    if (!c) return failc(RLG_ERR_INVALID, "out of host memory");
    c->e = e; c->cfg = *cfg;
    if (const char* ev = getenv("RLG_COLLECT_OVERLAP")) c->overlap = atoi(ev) != 0;
    if (const char* ev = getenv("RLG_COLLECT_CHAIN")) c->chain = atoi(ev) != 0;
    c->device = rlg_engine_device(e);
    c->A = rlg_engine_num_arenas(e); c->P = rlg_engine_num_players(e); c->N = c->A * c->P; c->obs = rlg_engine_obs_size(e);
    c->maxT = cfg->max_steps;
    c->rowBase = (uint64_t)rlg_engine_arena_id_base(e) * (uint64_t)c->P;  // sampling streams keyed by GLOBAL row id
    c->numLayers = cfg->num_hidden + 1;
    const int numActions = rlg_engine_num_actions(e);  // ActionParser::GetActionAmount of the engine's action table
    if (numActions < 1 || numActions > 96) { delete c; return failc(RLG_ERR_INVALID, "action head wider than 96"); }
    for (int n = 0; n < 2; n++) {
        int in = c->obs;
        for (int l = 0; l < c->numLayers; l++) {
            auto& h = c->L[n][l];
            h.in = in;
            h.out = (l < cfg->num_hidden) ? (n == 0 ? cfg->policy_hidden[l] : cfg->critic_hidden[l]) : (n == 0 ? numActions : 1);
            h.kPad = (in + kBlockK - 1) / kBlockK * kBlockK;
            h.nPad = (l < cfg->num_hidden) ? h.out : (h.out + 15) / 16 * 16;
            if (h.kPad > kMaxWidth) { delete c; return failc(RLG_ERR_INVALID, "observation wider than 256 floats is not supported by the fused MLP"); }
            in = h.out;
        }
    }
#define CKX(expr)                                                                                                                      \
    do {                                                                                                                               \
        cudaError_t _e = (expr);                                                                                                       \
        if (_e != cudaSuccess) { std::string m = std::string(#expr) + ": " + cudaGetErrorString(_e); rlg_collector_destroy(c); return failc(RLG_ERR_CUDA, m); } \
    } while (0)
    CKX(cudaSetDevice(c->device));
    for (int n = 0; n < 2; n++)
        for (int l = 0; l < c->numLayers; l++) {
            auto& h = c->L[n][l];
            CKX(cudaMalloc(&h.dW, (size_t)h.kPad * h.nPad * 4));
            CKX(cudaMalloc(&h.dB, (size_t)h.nPad * 4));
        }
    const size_t T = c->maxT, N = c->N;
    CKX(cudaMalloc(&c->dObs, (T + 1) * N * c->obs * 4));
    CKX(cudaMalloc(&c->dAction, T * N * 4));
    CKX(cudaMalloc(&c->dLogprob, T * N * 4));
    CKX(cudaMalloc(&c->dReward, T * N * 4));
    CKX(cudaMalloc(&c->dDone, T * c->A));
    CKX(cudaMalloc(&c->dValue, (T + 1) * N * 4));
    CKX(cudaMalloc(&c->dAdv, T * N * 4));
    CKX(cudaMalloc(&c->dTarget, T * N * 4));
    CKX(cudaMalloc(&c->dRet, T * N * 4));
    CKX(cudaFuncSetAttribute(k_mlp_infer, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    *out = c;
    return RLG_OK;
}

int rlg_collector_set_layer(rlg_collector* c, int net, int layer, const float* W, const float* b, int out_dim, int in_dim) {
    if (!c || !W || !b) return failc(RLG_ERR_INVALID, "null argument");
    if (net < 0 || net > 1 || layer < 0 || layer >= c->numLayers) return failc(RLG_ERR_INVALID, "bad net/layer index");
    auto& h = c->L[net][layer];
    if (out_dim != h.out || in_dim != h.in) return failc(RLG_ERR_INVALID, "layer shape does not match the collector configuration");
    CKC(cudaSetDevice(c->device));
    // pack into [kPad/32] canonical (nPad x 32) blocks, TF32-rounded, zero padded
    std::vector<float> pw((size_t)h.kPad * h.nPad, 0.f), pb((size_t)h.nPad, 0.f);
    for (int n = 0; n < h.out; n++) {
        for (int k = 0; k < h.in; k++) {
            size_t blk = (size_t)(k / kBlockK) * h.nPad * kBlockK;
            pw[blk + canon_off(n, k % kBlockK) / 4] = round_tf32_host(W[(size_t)n * h.in + k]);
        }
        pb[n] = b[n];
    }
    cudaStream_t s = (cudaStream_t)rlg_engine_stream(c->e);
    CKC(cudaMemcpyAsync(h.dW, pw.data(), pw.size() * 4, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(h.dB, pb.data(), pb.size() * 4, cudaMemcpyHostToDevice, s));
    CKC(cudaStreamSynchronize(s));
    h.set = true;
    return RLG_OK;
}

// The same from DEVICE memory (W_dev [out, in] row-major with leading dimension ldw, b_dev [out]): packed by a kernel on `stream`
// (NULL = the engine's stream), no host round trip — how the native PPO learner (ppo.cu) hands new weights to the agents.
int rlg_collector_set_layer_device(rlg_collector* c, int net, int layer, const float* W_dev, int ldw, const float* b_dev, int out_dim, int in_dim,
                                   void* stream) {
    if (!c || !W_dev || !b_dev) return failc(RLG_ERR_INVALID, "null argument");
    if (net < 0 || net > 1 || layer < 0 || layer >= c->numLayers) return failc(RLG_ERR_INVALID, "bad net/layer index");
    auto& h = c->L[net][layer];
    if (out_dim != h.out || in_dim != h.in || ldw < in_dim) return failc(RLG_ERR_INVALID, "layer shape does not match the collector configuration");
    CKC(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : (cudaStream_t)rlg_engine_stream(c->e);
    const int total = h.kPad * h.nPad;
    k_pack_layer<<<(total + 255) / 256, 256, 0, s>>>(W_dev, ldw, b_dev, h.out, h.in, h.kPad, h.nPad, h.dW, h.dB);
    c->launches++;
    CKC(cudaGetLastError());
    h.set = true;
    return RLG_OK;
}

int rlg_collector_infer(rlg_collector* c, const float* obs_dev, int n_rows, uint64_t counter, int32_t* action_dev, float* logprob_dev,
                        float* value_dev, void* stream) {
    if (!c || !obs_dev || n_rows < 0) return failc(RLG_ERR_INVALID, "bad argument");
    if (n_rows == 0) return RLG_OK;
    CKC(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : (cudaStream_t)rlg_engine_stream(c->e);
    return launch_infer(c, obs_dev, n_rows, counter, action_dev, logprob_dev, value_dev, s);
}

int rlg_collector_collect(rlg_collector* c, int n_steps, void* stream) {
    if (!c || n_steps < 1 || n_steps > c->maxT) return failc(RLG_ERR_INVALID, "n_steps must be in [1, max_steps]");
    { MlpNet tmp; for (int n = 0; n < 2; n++) { int rc = fill_net(c, n, tmp); if (rc != RLG_OK) return rc; } }
    CKC(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : (cudaStream_t)rlg_engine_stream(c->e);
    const size_t N = c->N, rowBytes = (size_t)c->obs * 4;
    if (!c->haveObs0) {
        float* eobs = nullptr;
        if (rlg_engine_outputs(c->e, &eobs, nullptr, nullptr) != RLG_OK) return failc(RLG_ERR_STATE, rlg_last_error());
        CKC(cudaMemcpyAsync(c->dObs, eobs, N * rowBytes, cudaMemcpyDeviceToDevice, s));
        c->haveObs0 = true;
    } else {
        // the obs after the previous collect's last step is this collect's first policy input
        CKC(cudaMemcpyAsync(c->dObs, c->dObs + (size_t)c->T * N * c->obs, N * rowBytes, cudaMemcpyDeviceToDevice, s));
    }
    c->T = n_steps;
    c->nStepEv = c->nInferEv = 0;
    auto mark = [&](std::vector<cudaEvent_t>& v, int& n) {
        if (!c->timing) return;
        if ((int)v.size() <= n) { cudaEvent_t ev; cudaEventCreate(&ev); v.push_back(ev); }
        cudaEventRecord(v[n++], s);
    };
    bool prevStepFused = false;  // the previous op on the stream is the fused step that wrote this inference's obs
    for (int t = 0; t < n_steps; t++) {
        mark(c->evInfer, c->nInferEv);
        int rc = launch_infer(c, c->dObs + (size_t)t * N * c->obs, (int)N, c->stepCounter, c->dAction + (size_t)t * N, c->dLogprob + (size_t)t * N,
                              c->dValue + (size_t)t * N, s, t > 0 && prevStepFused, c->chain && !c->stepHook && !c->resetHook);
        mark(c->evInfer, c->nInferEv);
        if (rc != RLG_OK) return rc;
        mark(c->evStep, c->nStepEv);
        if (c->stepHook) {  // host plugins finish the step (and re-set finished arenas themselves)
            if (rlg_engine_step_begin(c->e, c->dAction + (size_t)t * N, c->dObs + (size_t)(t + 1) * N * c->obs, c->dReward + (size_t)t * N,
                                      c->dDone + (size_t)t * c->A, s) != RLG_OK)
                return failc(RLG_ERR_STATE, rlg_last_error());
            if (c->stepHook(c->stepUser, t, c->dAction + (size_t)t * N, c->dObs + (size_t)(t + 1) * N * c->obs, c->dReward + (size_t)t * N,
                            c->dDone + (size_t)t * c->A) != RLG_OK)
                return failc(RLG_ERR_STATE, std::string("step hook failed: ") + rlg_last_error());
            mark(c->evStep, c->nStepEv);
            c->launches++;
            c->stepCounter++;
            prevStepFused = false;
            continue;
        }
        // the step as the programmatic dependent of the inference that produces its actions: every role block waits for the inference tiles
        // that cover its arenas' rows instead of for the whole inference, so a block never waits for another block's slow step (per-block
        // chain step -> inference tiles -> step ...; the blocks only meet again at the end of the collect)
        // Measured (profiles/r02x_collect_overlap_ab.txt): the Learner's collection gets 3 % faster, the bench loop (L2 flushed before every
        // collect) 6 % slower than with the inference overlap alone, so the chain is an opt-in (RLG_COLLECT_CHAIN=1).
        const bool chained = c->chain && c->overlap && !c->timing && !c->resetHook && c->dTileDone;
        if (rlg_engine_step_to_after(c->e, c->dAction + (size_t)t * N, c->dObs + (size_t)(t + 1) * N * c->obs, c->dReward + (size_t)t * N,
                                     c->dDone + (size_t)t * c->A, s, chained ? c->dTileDone : nullptr, c->tileSeq, kTileM) != RLG_OK)
            return failc(RLG_ERR_STATE, rlg_last_error());
        mark(c->evStep, c->nStepEv);
        c->launches++;
        c->stepCounter++;
        prevStepFused = !c->resetHook;
        if (c->resetHook) {  // host StateSetter: re-set finished arenas before the next inference reads their obs
            c->hDone.resize(c->A);
            CKC(cudaMemcpyAsync(c->hDone.data(), c->dDone + (size_t)t * c->A, c->A, cudaMemcpyDeviceToHost, s));
            CKC(cudaStreamSynchronize(s));
            c->hIds.clear();
            for (int a = 0; a < c->A; a++) if (c->hDone[a]) c->hIds.push_back(a);
            if (!c->hIds.empty()) c->resetHook(c->resetUser, c->hIds.data(), (int)c->hIds.size(), c->dObs + (size_t)(t + 1) * N * c->obs);
        }
    }
    // value of the state after the last step (Learner.cpp:618-640 appends nextStates[count-1])
    mark(c->evInfer, c->nInferEv);
    int rc = launch_infer(c, c->dObs + (size_t)n_steps * N * c->obs, (int)N, c->stepCounter, nullptr, nullptr, c->dValue + (size_t)n_steps * N, s, prevStepFused);
    mark(c->evInfer, c->nInferEv);
    return rc;
}

int rlg_collector_set_reset_hook(rlg_collector* c, rlg_reset_hook hook, void* user) {
    if (!c) return failc(RLG_ERR_INVALID, "null collector");
    c->resetHook = hook; c->resetUser = user;
    return RLG_OK;
}

// A trajectory collected elsewhere (host arrays, T-major like the ring) becomes the collector's "last collect": GAE, return
// statistics, export and the learner hand-over then work on it as on its own.  obs [T+1][N][obs], action [T][N] i32,
// logprob / reward [T][N], done [T][A] u8, value [T+1][N].
int rlg_collector_load_external(rlg_collector* c, int n_steps, const float* obs_host, const int32_t* action_host, const float* logprob_host,
                                const float* reward_host, const uint8_t* done_host, const float* value_host) {
    if (!c || !obs_host || !action_host || !logprob_host || !reward_host || !done_host || !value_host) return failc(RLG_ERR_INVALID, "null argument");
    if (n_steps < 1 || n_steps > c->maxT) return failc(RLG_ERR_INVALID, "n_steps must be in [1, max_steps]");
    CKC(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)rlg_engine_stream(c->e);
    const size_t T = n_steps, N = c->N;
    CKC(cudaMemcpyAsync(c->dObs, obs_host, (T + 1) * N * c->obs * 4, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(c->dAction, action_host, T * N * 4, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(c->dLogprob, logprob_host, T * N * 4, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(c->dReward, reward_host, T * N * 4, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(c->dDone, done_host, T * c->A, cudaMemcpyHostToDevice, s));
    CKC(cudaMemcpyAsync(c->dValue, value_host, (T + 1) * N * 4, cudaMemcpyHostToDevice, s));
    CKC(cudaStreamSynchronize(s));
    c->T = n_steps;
    c->haveObs0 = true;
    return RLG_OK;
}

int rlg_collector_set_step_hook(rlg_collector* c, rlg_step_hook hook, void* user) {
    if (!c) return failc(RLG_ERR_INVALID, "null collector");
    c->stepHook = hook; c->stepUser = user;
    return RLG_OK;
}

int rlg_collector_enable_timing(rlg_collector* c, int on) {
    if (!c) return failc(RLG_ERR_INVALID, "null collector");
    c->timing = on != 0;
    return RLG_OK;
}

int rlg_collector_kernel_times(rlg_collector* c, double* step_ms, int32_t* step_launches, double* infer_ms, int32_t* infer_launches) {
    if (!c) return failc(RLG_ERR_INVALID, "null collector");
    CKC(cudaSetDevice(c->device));
    double st = 0, in = 0;
    for (int i = 0; i + 1 < c->nStepEv; i += 2) {
        CKC(cudaEventSynchronize(c->evStep[i + 1]));
        float ms = 0; CKC(cudaEventElapsedTime(&ms, c->evStep[i], c->evStep[i + 1])); st += ms;
    }
    for (int i = 0; i + 1 < c->nInferEv; i += 2) {
        CKC(cudaEventSynchronize(c->evInfer[i + 1]));
        float ms = 0; CKC(cudaEventElapsedTime(&ms, c->evInfer[i], c->evInfer[i + 1])); in += ms;
    }
    if (step_ms) *step_ms = st;
    if (step_launches) *step_launches = c->nStepEv / 2;
    if (infer_ms) *infer_ms = in;
    if (infer_launches) *infer_launches = c->nInferEv / 2;
    return RLG_OK;
}

int rlg_collector_gae(rlg_collector* c, float gamma, float lambda, float return_std, float clip_range, void* stream) {
    if (!c) return failc(RLG_ERR_INVALID, "null collector");
    if (c->T < 1) return failc(RLG_ERR_STATE, "rlg_collector_collect has not run");
    CKC(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : (cudaStream_t)rlg_engine_stream(c->e);
    k_gae<<<(c->N + 127) / 128, 128, 0, s>>>(c->T, c->N, c->P, c->dReward, c->dDone, c->dValue, c->dAdv, c->dTarget, c->dRet, gamma, lambda, return_std,
                                             clip_range);
    c->launches++;
    CKC(cudaGetLastError());
    return RLG_OK;
}

int rlg_collector_view(rlg_collector* c, rlg_traj_view* o) {
    if (!c || !o) return failc(RLG_ERR_INVALID, "null argument");
    o->T = c->T; o->N = c->N; o->A = c->A; o->P = c->P; o->obs_size = c->obs;
    o->obs = c->dObs; o->action = c->dAction; o->logprob = c->dLogprob; o->reward = c->dReward; o->done = c->dDone;
    o->value = c->dValue; o->advantage = c->dAdv; o->value_target = c->dTarget; o->ret = c->dRet;
    return RLG_OK;
}

int rlg_collector_export(rlg_collector* c, float* states, int64_t* actions, float* log_probs, float* rewards, float* next_states, float* dones,
                         float* truncateds, float* value_targets, float* advantages, void* stream) {
    if (!c) return failc(RLG_ERR_INVALID, "null collector");
    if (c->T < 1) return failc(RLG_ERR_STATE, "rlg_collector_collect has not run");
    CKC(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : (cudaStream_t)rlg_engine_stream(c->e);
    long rows = (long)c->T * c->N;
    int block = 256;
    long grid = (rows * 32 + block - 1) / block;
    k_export_rows<<<(unsigned)grid, block, 0, s>>>(c->T, c->N, c->P, c->obs, c->dObs, c->dAction, c->dLogprob, c->dReward, c->dDone, c->dTarget, c->dAdv,
                                                   states, actions, log_probs, rewards, next_states, dones, truncateds, value_targets, advantages);
    c->launches++;
    CKC(cudaGetLastError());
    return RLG_OK;
}

int rlg_collector_return_stats(rlg_collector* c, double* out3_host, float* first_returns_host, int n_first, void* stream) {
    if (!c || !out3_host) return failc(RLG_ERR_INVALID, "null argument");
    if (c->T < 1) return failc(RLG_ERR_STATE, "rlg_collector_collect has not run");
    CKC(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : (cudaStream_t)rlg_engine_stream(c->e);
    const long total = (long)c->T * c->N;
    if (n_first < 0 || !first_returns_host) n_first = 0;
    if (n_first > total) n_first = (int)total;
    if (!c->dStats) CKC(cudaMalloc(&c->dStats, 3 * 8 + 4096 * 4));
    if (n_first > 4096) return failc(RLG_ERR_INVALID, "n_first must be <= 4096");
    float* first = reinterpret_cast<float*>(c->dStats + 3);
    CKC(cudaMemsetAsync(c->dStats, 0, 24, s));
    k_return_stats<<<148, 256, 0, s>>>(c->T, c->N, c->dRet, c->dAdv, c->dTarget, c->dStats, first, n_first);
    c->launches++;
    CKC(cudaGetLastError());
    CKC(cudaMemcpyAsync(out3_host, c->dStats, 24, cudaMemcpyDeviceToHost, s));
    if (n_first > 0) CKC(cudaMemcpyAsync(first_returns_host, first, (size_t)n_first * 4, cudaMemcpyDeviceToHost, s));
    CKC(cudaStreamSynchronize(s));
    for (int i = 0; i < 3; i++) out3_host[i] /= (double)total;
    return RLG_OK;
}

uint64_t rlg_collector_launch_count(const rlg_collector* c) { return c ? c->launches : 0; }

}  // extern "C"
