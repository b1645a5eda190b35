// gemm.cu — the dense contractions of the PPO minibatch update on the 5th-gen tensor cores.
//
// PPOLearner::Learn (P/private/RLGymPPO_CPP/PPO/PPOLearner.cpp:125-290) is, per minibatch and per network, three GEMMs
// per Linear layer: the forward Y = X W^T + b, and in the backward pass dX = dY W and dW = dY^T X.  All three are the
// same kernel here:
//
//     C[M, N] (+)= A[M, K] . B[N, K]^T (+ bias[N]) (ReLU)          A, B, C row-major fp32, both operands K-major
//
//   forward : A = X [rows, in],    B = W   [out, in]                     M = rows, N = out, K = in
//   dX      : A = dY [rows, out],  B = W^T [in, out]                     M = rows, N = in,  K = out
//   dW      : A = dY^T [out, rows], B = X^T [in, rows]  split over K     M = out,  N = in,  K = rows   (atomic accumulate)
//
// One CTA = one 128 x NT accumulator tile in TMEM (NT <= 256 columns), 256 threads.  Per 32-wide K block every thread
// moves float4s from global memory into shared memory in the UMMA canonical K-major no-swizzle layout (8-row x 16-byte
// core matrices), rounding to TF32; two stages, recycled through mbarriers that tcgen05.commit arrives on; one thread
// issues tcgen05.mma kind::tf32 (4 x K=8 per block).  The epilogue reads the accumulator with tcgen05.ld (both
// warpgroups, alternate 32-column chunks), applies bias / ReLU and stores or atomically adds (split-K) into C.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <string>

#include "../../include/rlgym_b200.h"
#include "pdl.h"

extern "C" void rlg_internal_set_error(const char* msg);

namespace {

constexpr int kGM = 128, kGK = 32, kGThreads = 256, kGStages = 2;
constexpr int kGABytes = kGM * kGK * 4;        // 16 KB per stage
constexpr int kGBBytes = 256 * kGK * 4;        // 32 KB per stage (NT <= 256)
constexpr int kGBar = kGStages * (kGABytes + kGBBytes);
constexpr int kGSmem = kGBar + 64;

struct GemmArgs {
    const float* A; const float* B; float* C; const float* bias;
    int32_t M, N, K, lda, ldb, ldc;
    int32_t flags;    // RLG_GEMM_*
    int32_t kbPerSplit;  // K blocks per blockIdx.z
    const float* mask; int32_t ldm;  // optional [M, N]: C = mask > 0 ? C : 0 (ReLU backward with the layer's output)
    float* Ct; int32_t ldct;         // optional [N, M]: the transposed result as well (the K-major operand of a later dW GEMM)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
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
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// round-to-nearest (ties away) to TF32 like cvt.rna.tf32.f32, on the integer pipe
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE, version 1 (Blackwell): LBO = 128 B, SBO = 1024 B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    const uint64_t lbo = 128 >> 4, sbo = 1024 >> 4;
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// One (rows x 32) K block of a row-major matrix on its way to canonical K-major shared memory, in two halves so that the
// global loads of the NEXT block are in flight while the tensor core works on the current one: slot idx = t + u * 256
// covers (row = idx[2:0] | idx[..:6] << 3, float4 column = idx[5:3]): the 8 lanes of a 128-bit shared-memory store phase
// hold 8 consecutive rows of one 16-byte column = one contiguous 128-byte core matrix (conflict free; the row-major
// mapping made every phase an 8-way bank conflict); rows beyond `rowsValid` and k beyond K read as 0.
template <int U>
__device__ __forceinline__ void block_load(float4 (&v)[U], const float* __restrict__ src, int ld, int row0, int rowsValid, int rowsTile, int k0, int K, int t) {
    const int total = rowsTile * (kGK / 4);
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int idx = t + u * kGThreads;
        const int r = (idx & 7) | ((idx >> 6) << 3), k = k0 + ((idx >> 3) & 7) * 4;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < total && row0 + r < rowsValid && k < K) v[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * ld + k));
    }
}
template <int U>
__device__ __forceinline__ void block_store(uint8_t* dst, const float4 (&v)[U], int rowsTile, int t) {
    const int total = rowsTile * (kGK / 4);
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int idx = t + u * kGThreads;
        if (idx < total) {
            const uint32_t r = (uint32_t)((idx & 7) | ((idx >> 6) << 3)), c4 = (uint32_t)((idx >> 3) & 7);
            float4 o = make_float4(to_tf32(v[u].x), to_tf32(v[u].y), to_tf32(v[u].z), to_tf32(v[u].w));
            *reinterpret_cast<float4*>(dst + (r >> 3) * 1024u + c4 * 128u + (r & 7u) * 16u) = o;
        }
    }
}

__global__ void __launch_bounds__(kGThreads, 1) k_gemm_tf32(const GemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGBar);  // [0..stages): stage free (its MMAs completed)
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(smem + kGBar + 32);
    const int t = threadIdx.x, warp = t >> 5;
    const int m0 = blockIdx.x * kGM;
    const int n0 = blockIdx.y * 256;
    const int nValid = g.N - n0 < 256 ? g.N - n0 : 256;
    const int NT = (nValid + 15) & ~15;                       // UMMA N: multiple of 16 for M = 128
    const uint32_t tmemCols = NT <= 32 ? 32u : (NT <= 64 ? 64u : (NT <= 128 ? 128u : 256u));
    const int nkbAll = (g.K + kGK - 1) / kGK;
    const int kbBegin = blockIdx.z * g.kbPerSplit;
    const int kbEnd = kbBegin + g.kbPerSplit < nkbAll ? kbBegin + g.kbPerSplit : nkbAll;
    const int nIt = kbEnd - kbBegin;
    if (nIt <= 0) return;  // uniform per CTA
    const uint32_t barBase = smem_u32(&bars[0]);

    if (t == 0) {
        for (int i = 0; i < kGStages; i++) mbar_init(barBase + 8 * i, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmemSlot)), "r"(tmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmemBase = *tmemSlot;
    const uint32_t idesc = make_idesc(kGM, NT);
    pdl_enter();  // (see k_gemm_tma)

    constexpr int kUA = kGM * (kGK / 4) / kGThreads;   // 4 float4 per thread for the A block
    constexpr int kUB = 256 * (kGK / 4) / kGThreads;   // 8 for a full-width B block
    float4 ra[kUA], rb[kUB];
    block_load<kUA>(ra, g.A, g.lda, m0, g.M, kGM, kbBegin * kGK, g.K, t);
    block_load<kUB>(rb, g.B, g.ldb, n0, g.N, NT, kbBegin * kGK, g.K, t);
    for (int it = 0; it < nIt; it++) {
        const int s = it & 1;
        uint8_t* sA = smem + s * (kGABytes + kGBBytes);
        uint8_t* sB = sA + kGABytes;
        if (it >= kGStages) mbar_wait(barBase + 8 * s, ((it >> 1) - 1) & 1);  // the MMAs that read this stage two iterations ago are done
        block_store<kUA>(sA, ra, kGM, t);
        block_store<kUB>(sB, rb, NT, t);
        fence_proxy_async();
        __syncthreads();
        if (it + 1 < nIt) {  // next block's loads fly while this block's MMAs run
            const int k0 = (kbBegin + it + 1) * kGK;
            block_load<kUA>(ra, g.A, g.lda, m0, g.M, kGM, k0, g.K, t);
            block_load<kUB>(rb, g.B, g.ldb, n0, g.N, NT, k0, g.K, t);
        }
        if (t == 0) {
            tc_fence_after();
            const uint32_t aBase = smem_u32(sA), bBase = smem_u32(sB);
#pragma unroll
            for (int j = 0; j < kGK / 8; j++) tc_mma_tf32(tmemBase, make_desc(aBase + j * 256), make_desc(bBase + j * 256), idesc, (it > 0 || j > 0) ? 1u : 0u);
            tc_commit(barBase + 8 * s);
        }
    }
    {   // commits complete in order: the last stage's barrier covers every MMA
        const int last = nIt - 1;
        mbar_wait(barBase + 8 * (last & 1), (last >> 1) & 1);
        tc_fence_after();
    }

    // epilogue: thread owns accumulator row (t & 127); the two warpgroups take alternate 32-column chunks
    const int r = t & (kGM - 1), half = t >> 7;
    const uint32_t tmemLane = tmemBase + ((uint32_t)((warp & 3) * 32) << 16);
    const int row = m0 + r;
    const bool addBias = g.bias != nullptr && blockIdx.z == 0;
    for (int c = half; c * 32 < NT; c += 2) {
        uint32_t v[32];
        tc_ld32(tmemLane + c * 32, v);
        tc_wait_ld();
        if (row < g.M) {
            float* dst = g.C + (size_t)row * g.ldc + n0 + c * 32;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int n = c * 32 + 4 * q;
                if (n >= nValid) break;
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    o[i] = __uint_as_float(v[4 * q + i]);
                    if (addBias && n + i < nValid) o[i] += __ldg(g.bias + n0 + n + i);
                    if (g.flags & RLG_GEMM_RELU) o[i] = fmaxf(o[i], 0.f);
                    if (g.mask && n + i < nValid && !(__ldg(g.mask + (size_t)row * g.ldm + n0 + n + i) > 0.f)) o[i] = 0.f;
                }
                if (g.Ct) {  // lanes = consecutive rows: one 128-byte line per column
#pragma unroll
                    for (int i = 0; i < 4; i++) if (n + i < nValid) g.Ct[(size_t)(n0 + n + i) * g.ldct + row] = o[i];
                }
                if (g.flags & RLG_GEMM_ATOMIC) {
                    if (n + 3 < nValid && (g.flags & RLG_GEMM_SCALAR_STORE) == 0) {
                        atomicAdd(reinterpret_cast<float4*>(dst + 4 * q), make_float4(o[0], o[1], o[2], o[3]));  // one 16-byte reduction
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; i++) if (n + i < nValid) atomicAdd(dst + 4 * q + i, o[i]);
                    }
                } else if (n + 3 < nValid && (g.flags & (RLG_GEMM_ACCUMULATE | RLG_GEMM_SCALAR_STORE)) == 0) {
                    *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) if (n + i < nValid) dst[4 * q + i] = ((g.flags & RLG_GEMM_ACCUMULATE) ? dst[4 * q + i] : 0.f) + o[i];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(tmemCols) : "memory");
}


// ---- the same GEMM fed by the TMA ---------------------------------------------------------------------------------------
// Operand blocks (128 x 32 of A, NT x 32 of B) arrive as cp.async.bulk.tensor boxes in the 128-byte swizzled K-major layout
// (one 128-byte row per matrix row, 16-byte chunks XOR-ed with row % 8 — conflict free for the tensor core, nothing passes
// through registers), S stages recycled through full / empty mbarriers: one producer thread, one MMA-issuing thread, both
// warpgroups in the epilogue.  The tensor core reads the fp32 bits as TF32 (low mantissa bits ignored).
struct GemmTmaArgs {
    GemmArgs g;
    int32_t stages, stageBytes;
};
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map),
                 "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: SBO = 1024 B (8 rows x 128 B), LBO unused, version 1
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(kGThreads, 1) k_gemm_tma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                           const GemmTmaArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const GemmArgs& g = a.g;
    const int S = a.stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * a.stageBytes);  // full[S], empty[S], done
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(smem + S * a.stageBytes + 8 * (2 * 4 + 1));
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int m0 = blockIdx.x * kGM;
    const int n0 = blockIdx.y * 256;
    const int nValid = g.N - n0 < 256 ? g.N - n0 : 256;
    const int NT = (nValid + 15) & ~15;
    const uint32_t tmemCols = NT <= 32 ? 32u : (NT <= 64 ? 64u : (NT <= 128 ? 128u : 256u));
    const int nkbAll = (g.K + kGK - 1) / kGK;
    const int kbBegin = blockIdx.z * g.kbPerSplit;
    const int kbEnd = kbBegin + g.kbPerSplit < nkbAll ? kbBegin + g.kbPerSplit : nkbAll;
    const int nIt = kbEnd - kbBegin;
    if (nIt <= 0) return;
    const uint32_t barFull = smem_u32(&bars[0]), barEmpty = smem_u32(&bars[4]), barDone = smem_u32(&bars[8]);

    if (t == 0) {
        for (int i = 0; i < S; i++) { mbar_init(barFull + 8 * i, 1); mbar_init(barEmpty + 8 * i, 1); }
        mbar_init(barDone, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmemSlot)), "r"(tmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmemBase = *tmemSlot;
    pdl_enter();  // barriers and TMEM are set up: from here on the operands (written by the kernels before this one) are read

    if (warp == 0 && lane == 0) {  // producer
        // a TMA box always delivers its full size (out-of-bounds rows / columns arrive as zeros): A box + B box = one stage
        const uint32_t bytesBox = (uint32_t)a.stageBytes;
        for (int it = 0; it < nIt; it++) {
            const int s = it % S, k = it / S;
            if (k >= 1) mbar_wait(barEmpty + 8 * s, (k - 1) & 1);
            mbar_expect_tx(barFull + 8 * s, bytesBox);
            const uint32_t dst = smem_u32(smem + s * a.stageBytes);
            const int k0 = (kbBegin + it) * kGK;
            tma_load_2d(dst, &mapA, k0, m0, barFull + 8 * s);
            tma_load_2d(dst + kGABytes, &mapB, k0, n0, barFull + 8 * s);
        }
    } else if (warp == 1 && lane == 0) {  // MMA issuer
        const uint32_t idesc = make_idesc(kGM, NT);
        for (int it = 0; it < nIt; it++) {
            const int s = it % S, k = it / S;
            mbar_wait(barFull + 8 * s, k & 1);
            tc_fence_after();
            const uint32_t aBase = smem_u32(smem + s * a.stageBytes), bBase = aBase + kGABytes;
#pragma unroll
            for (int j = 0; j < kGK / 8; j++)
                tc_mma_tf32(tmemBase, make_desc_sw128(aBase + j * 32), make_desc_sw128(bBase + j * 32), idesc, (it > 0 || j > 0) ? 1u : 0u);
            tc_commit(barEmpty + 8 * s);
        }
        tc_commit(barDone);
    }
    __syncwarp();
    mbar_wait(barDone, 0);
    tc_fence_after();

    // epilogue: thread t owns accumulator row t & 127 (TMEM lane), the two warpgroups take alternate 32-column chunks.  Per chunk the
    // row's 32 mask values (ReLU backward) are fetched as eight independent 16-byte loads BEFORE the accumulator load is waited
    // for, the bias comes in as float4, C goes out as 16-byte stores and the transposed copy as 32 stores that are coalesced across
    // the warp (consecutive lanes = consecutive rows = consecutive addresses of one C^T row).
    const int r = t & (kGM - 1), half = t >> 7;
    const uint32_t tmemLane = tmemBase + ((uint32_t)((warp & 3) * 32) << 16);
    const int row = m0 + r;
    const bool rowOk = row < g.M;
    const bool addBias = g.bias != nullptr && blockIdx.z == 0;
    const bool vecMask = g.mask != nullptr && (g.ldm & 3) == 0 && (((uintptr_t)g.mask) & 15) == 0;
    const bool vecBias = addBias && (((uintptr_t)g.bias) & 15) == 0;
    for (int c = half; c * 32 < NT; c += 2) {
        uint32_t v[32];
        tc_ld32(tmemLane + c * 32, v);
        const int nc = n0 + c * 32;                                  // first global column of the chunk
        const int valid = nValid - c * 32 < 32 ? nValid - c * 32 : 32;  // columns of the chunk inside N
        float4 mk[8];
        if (g.mask && rowOk) {
            const float* mrow = g.mask + (size_t)row * g.ldm + nc;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (vecMask && 4 * q + 3 < valid) mk[q] = __ldg(reinterpret_cast<const float4*>(mrow + 4 * q));
                else {
                    mk[q].x = 4 * q + 0 < valid ? __ldg(mrow + 4 * q + 0) : 1.f; mk[q].y = 4 * q + 1 < valid ? __ldg(mrow + 4 * q + 1) : 1.f;
                    mk[q].z = 4 * q + 2 < valid ? __ldg(mrow + 4 * q + 2) : 1.f; mk[q].w = 4 * q + 3 < valid ? __ldg(mrow + 4 * q + 3) : 1.f;
                }
            }
        }
        tc_wait_ld();
        if (!rowOk) continue;
        float o[32];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (addBias) {
                if (vecBias && 4 * q + 3 < valid) b = __ldg(reinterpret_cast<const float4*>(g.bias + nc + 4 * q));
                else {
                    if (4 * q + 0 < valid) b.x = __ldg(g.bias + nc + 4 * q + 0);
                    if (4 * q + 1 < valid) b.y = __ldg(g.bias + nc + 4 * q + 1);
                    if (4 * q + 2 < valid) b.z = __ldg(g.bias + nc + 4 * q + 2);
                    if (4 * q + 3 < valid) b.w = __ldg(g.bias + nc + 4 * q + 3);
                }
            }
            o[4 * q + 0] = __uint_as_float(v[4 * q + 0]) + b.x; o[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + b.y;
            o[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + b.z; o[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + b.w;
        }
        if (g.flags & RLG_GEMM_RELU) {
#pragma unroll
            for (int i = 0; i < 32; i++) o[i] = fmaxf(o[i], 0.f);
        }
        if (g.mask) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (!(mk[q].x > 0.f)) o[4 * q + 0] = 0.f;
                if (!(mk[q].y > 0.f)) o[4 * q + 1] = 0.f;
                if (!(mk[q].z > 0.f)) o[4 * q + 2] = 0.f;
                if (!(mk[q].w > 0.f)) o[4 * q + 3] = 0.f;
            }
        }
        if (g.Ct) {
            float* ct = g.Ct + (size_t)nc * g.ldct + row;
#pragma unroll
            for (int i = 0; i < 32; i++) if (i < valid) ct[(size_t)i * g.ldct] = o[i];
        }
        float* dst = g.C + (size_t)row * g.ldc + nc;
        if (g.flags & RLG_GEMM_ATOMIC) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (4 * q + 3 < valid && (g.flags & RLG_GEMM_SCALAR_STORE) == 0) {
                    atomicAdd(reinterpret_cast<float4*>(dst + 4 * q), make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));  // one 16-byte reduction
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) if (4 * q + i < valid) atomicAdd(dst + 4 * q + i, o[4 * q + i]);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (4 * q + 3 < valid && (g.flags & (RLG_GEMM_ACCUMULATE | RLG_GEMM_SCALAR_STORE)) == 0) {
                    *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (4 * q + i < valid) dst[4 * q + i] = ((g.flags & RLG_GEMM_ACCUMULATE) ? dst[4 * q + i] : 0.f) + o[4 * q + i];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(tmemCols) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
bool g_encode_tried = false;

// row-major [rows, K] fp32 matrix, boxes of 32 K-elements (128 B) x boxRows rows, 128-byte swizzle, zero fill out of bounds
bool make_map(CUtensorMap* map, const float* base, int rows, int K, int ld, int boxRows) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)kGK, (cuuint32_t)boxRows};
    cuuint32_t estr[2] = {1, 1};
    return g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool g_attr_tma[64] = {};

int failg(int code, const std::string& m) { rlg_internal_set_error(m.c_str()); return code; }
bool g_attr_set[64] = {};

}  // namespace

extern "C" int rlg_gemm_tf32_fused(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias,
                                   int flags, int split_k, const float* mask, int ldm, float* Ct, int ldct, void* stream);
extern "C" int rlg_gemm_tf32(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias, int flags,
                             int split_k, void* stream) {
    return rlg_gemm_tf32_fused(M, N, K, A, lda, B, ldb, C, ldc, bias, flags, split_k, nullptr, 0, nullptr, 0, stream);
}
extern "C" int rlg_gemm_tf32_fused(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias,
                                   int flags, int split_k, const float* mask, int ldm, float* Ct, int ldct, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0 || !A || !B || !C) return failg(RLG_ERR_INVALID, "rlg_gemm_tf32: bad argument");
    if ((K & 3) || (lda & 3) || (ldb & 3) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15))
        return failg(RLG_ERR_INVALID, "rlg_gemm_tf32: K, lda and ldb must be multiples of 4 floats and A, B 16-byte aligned");
    if ((ldc & 3) || ((uintptr_t)C & 15)) flags |= RLG_GEMM_SCALAR_STORE;
    if (split_k < 1) split_k = 1;
    if (split_k > 1 && !(flags & RLG_GEMM_ATOMIC)) return failg(RLG_ERR_INVALID, "rlg_gemm_tf32: split_k > 1 needs RLG_GEMM_ATOMIC (C accumulates)");
    if ((flags & RLG_GEMM_ATOMIC) && ((flags & RLG_GEMM_RELU) || mask || Ct))
        return failg(RLG_ERR_INVALID, "rlg_gemm_tf32: ReLU / mask / transposed output cannot follow a partial sum");
    if ((mask && ldm < N) || (Ct && ldct < M)) return failg(RLG_ERR_INVALID, "rlg_gemm_tf32: bad mask / transposed-output leading dimension");
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return failg(RLG_ERR_CUDA, std::string("rlg_gemm_tf32: ") + cudaGetErrorString(err));
    if (dev >= 0 && dev < 64 && !g_attr_set[dev]) {
        err = cudaFuncSetAttribute(k_gemm_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmem);
        if (err != cudaSuccess) return failg(RLG_ERR_CUDA, std::string("rlg_gemm_tf32: ") + cudaGetErrorString(err));
        g_attr_set[dev] = true;
    }
    GemmArgs g;
    g.A = A; g.B = B; g.C = C; g.bias = bias; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.flags = flags;
    g.mask = mask; g.ldm = ldm; g.Ct = Ct; g.ldct = ldct;
    const int nkb = (K + kGK - 1) / kGK;
    if (split_k > nkb) split_k = nkb;
    g.kbPerSplit = (nkb + split_k - 1) / split_k;
    split_k = (nkb + g.kbPerSplit - 1) / g.kbPerSplit;
    dim3 grid((M + kGM - 1) / kGM, (N + 255) / 256, split_k);
    if (!g_encode_tried) {
        g_encode_tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            g_encode = (EncodeTiledFn)fn;
        if (const char* ev = getenv("RLG_GEMM_NO_TMA")) { if (atoi(ev)) g_encode = nullptr; }  // A/B switch: the register-staged kernel
    }
    if (g_encode) {
        const int boxRowsB = N >= 256 ? 256 : ((N + 15) & ~15);
        GemmTmaArgs ta;
        ta.g = g;
        ta.stageBytes = kGABytes + boxRowsB * kGK * 4;
        ta.stages = (100 * 1024) / ta.stageBytes;  // two CTAs per SM ...
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if ((long long)grid.x * grid.y * grid.z <= sms) ta.stages = 4;  // ... unless one wave of single CTAs (split-K): a deeper pipeline instead
        if (ta.stages > 4) ta.stages = 4;
        if (ta.stages < 2) ta.stages = 2;
        const int smemBytes = ta.stages * ta.stageBytes + 128;
        CUtensorMap mapA, mapB;
        if (!make_map(&mapA, A, M, K, lda, kGM) || !make_map(&mapB, B, N, K, ldb, boxRowsB)) return failg(RLG_ERR_CUDA, "rlg_gemm_tf32: cuTensorMapEncodeTiled failed");
        if (dev >= 0 && dev < 64 && !g_attr_tma[dev]) {
            err = cudaFuncSetAttribute(k_gemm_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (kGABytes + kGBBytes) + 128);
            if (err != cudaSuccess) return failg(RLG_ERR_CUDA, std::string("rlg_gemm_tf32: ") + cudaGetErrorString(err));
            g_attr_tma[dev] = true;
        }
        err = launch_pdl(k_gemm_tma, grid, dim3(kGThreads), (size_t)smemBytes, (cudaStream_t)stream, mapA, mapB, ta);
    } else {
        err = launch_pdl(k_gemm_tf32, grid, dim3(kGThreads), (size_t)kGSmem, (cudaStream_t)stream, g);
    }
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) return failg(RLG_ERR_CUDA, std::string("rlg_gemm_tf32 launch: ") + cudaGetErrorString(err));
    return RLG_OK;
}
