// ppo.cu — PPOLearner::Learn and the ExperienceBuffer on the device: no autograd, no torch.
//
// Reference: P/private/RLGymPPO_CPP/PPO/PPOLearner.cpp:67-349 (Learn: epochs x shuffled batches x minibatches, clipped PPO loss
// with entropy bonus, MSE value loss, clip_grad_norm_ 0.5, Adam), PPO/DiscretePolicy.cpp:64-75 (GetBackpropData: softmax of
// logits / temperature, clamp to [1e-11, 1], log, gather, entropy), PPO/ExperienceBuffer.cpp:12-121 (FIFO + shuffled batches).
//
// One minibatch = one pass over these kernels on one stream:
//   k_gather_rows      the shuffled rows of the experience ring -> X [n, obsP] and X^T [obsP, n] (both K-major operands of the
//                      first layer's forward and weight-gradient GEMMs) + their action / advantage / old log-prob / value target
//   rlg_gemm_tf32_fused  (csrc/gemm.cu, TMA + tcgen05 TF32)   every Linear layer's forward, input-gradient and weight-gradient GEMM
//   k_policy_loss      softmax -> clamp -> log-prob / entropy -> ratio / clipped surrogate -> dLogits and dLogits^T in one pass
//                      (forward AND backward of the loss; the SB3 diagnostics of PPOLearner.cpp:181-196 as block sums)
//   k_value_loss       MSE -> dV, dV^T
//   k_bias_grad        db += row sums of dY^T
// and once per batch: the all-reduce hook on the ONE flat gradient vector of both networks, k_grad_norm (per-network global norm),
// k_adam (clip-by-norm + Adam for every parameter of both networks in one launch, gradients zeroed for the next batch) and
// k_transpose (W^T copies, the K-major operand of the input-gradient GEMMs).
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>
#include <random>
#include <string>
#include <vector>

#include "../../include/rlgym_b200.h"
#include "pdl.h"

extern "C" void rlg_internal_set_error(const char* msg);

namespace {

constexpr int kMaxL = RLG_MAX_HIDDEN_LAYERS + 1;
constexpr float kMinProb = 1e-11f;  // DiscretePolicy::ACTION_MIN_PROB (DiscretePolicy.h:19)
constexpr int kAccWords = 12;       // entropy, kl, ratio, value loss, clip fraction, normSq[2], diffSq[2]

inline int pad4(int n) { return (n + 3) & ~3; }
int failp(int code, const std::string& m) { rlg_internal_set_error(m.c_str()); return code; }
#define CKP(expr)                                                                                             \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) return failp(RLG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
#define CKR(expr) do { int _rc = (expr); if (_rc != RLG_OK) return _rc; } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

// ---- ExperienceBuffer::GetAllBatchesShuffled: a fresh permutation = sort of random keys ------------------------------------------
__global__ void k_shuffle_keys(uint64_t* keys, int32_t* idx, long n, uint64_t seed, uint64_t counter) {
    pdl_enter();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = mix64(mix64(seed ^ (counter * 0xD1B54A32D192ED03ull)) + (uint64_t)i);
    idx[i] = (int32_t)i;
}

// ---- gather: 32 rows per block, transposed through shared memory ------------------------------------------------------------------
// logical row r of the FIFO lives at ring slot (head + r) % cap
__global__ void __launch_bounds__(256) k_gather_rows(const int32_t* __restrict__ perm, long n, long ldT, long head, long cap, int obs, int obsP,
                                                     const float* __restrict__ bStates, const int64_t* __restrict__ bActions,
                                                     const float* __restrict__ bLogp, const float* __restrict__ bTarget, const float* __restrict__ bAdv,
                                                     float* __restrict__ X, float* __restrict__ Xt, int32_t* __restrict__ act, float* __restrict__ oldLp,
                                                     float* __restrict__ tgt, float* __restrict__ adv) {
    extern __shared__ float tile[];  // [32][obsP + 1]
    pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long row0 = (long)blockIdx.x * 32;
    const int ts = obsP + 1;
    for (int r = warp; r < 32; r += 8) {
        const long row = row0 + r;
        if (row < n) {
            long src = head + perm[row];
            if (src >= cap) src -= cap;
            const float* s = bStates + (size_t)src * obs;
            for (int k = lane; k < obsP; k += 32) tile[r * ts + k] = k < obs ? s[k] : 0.f;
            if (lane == 0) { act[row] = (int32_t)bActions[src]; oldLp[row] = bLogp[src]; tgt[row] = bTarget[src]; adv[row] = bAdv[src]; }
        } else {
            for (int k = lane; k < obsP; k += 32) tile[r * ts + k] = 0.f;
        }
    }
    __syncthreads();
    const int rowsHere = (int)((n - row0) < 32 ? (n - row0) : 32);
    // X: the block's 32 rows are contiguous in [n, obsP]
    for (int i = threadIdx.x; i < rowsHere * obsP; i += 256) X[(size_t)row0 * obsP + i] = tile[(i / obsP) * ts + (i % obsP)];
    // X^T: one 128-byte line per feature
    for (int k = warp; k < obsP; k += 8)
        if (lane < rowsHere) Xt[(size_t)k * ldT + row0 + lane] = tile[lane * ts + k];
}

// ---- policy loss forward + backward ----------------------------------------------------------------------------------------------
// per row: p = softmax(z / T); pc = clamp(p, 1e-11, 1); lp = log pc; H = -sum lp pc; ratio = exp(lp[a] - old);
// L = ((-mean min(ratio adv, clamp(ratio, 1-c, 1+c) adv)) - entCoef mean H) * ratioB  (PPOLearner.cpp:139-178)
// dL/dz_j = (1/T) p_j (g_j - sum_k g_k p_k),  g_j = [p_j in clamp range] (entCoef ratioB / n (lp_j + 1) + [j = a] gLp / pc_a),
// gLp = -(ratioB / n) ratio adv [ratio in clip range or the unclipped surrogate is the smaller]   (torch.min / clamp subgradients)
__global__ void __launch_bounds__(256) k_policy_loss(const float* __restrict__ logits, int ldz, int nAct, int nActP, long n, long ldT,
                                                     const int32_t* __restrict__ act, const float* __restrict__ adv, const float* __restrict__ oldLp,
                                                     float invTemp, float clipRange, float entCoef, float ratioB, float* __restrict__ dZ,
                                                     float* __restrict__ dZt, double* __restrict__ acc) {
    extern __shared__ float tile[];  // [32][nActP + 1]
    pdl_enter();
    __shared__ float red[8][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long row0 = (long)blockIdx.x * 32;
    const int ts = nActP + 1;
    const float invN = 1.f / (float)n;
    float sEnt = 0.f, sKl = 0.f, sRatio = 0.f, sClip = 0.f;
    for (int r = warp; r < 32; r += 8) {
        const long row = row0 + r;
        if (row >= n) {
            for (int j = lane; j < nActP; j += 32) tile[r * ts + j] = 0.f;
            continue;
        }
        const float* z = logits + (size_t)row * ldz;
        float zz[4], p[4], lp[4];
        float mx = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = lane + 32 * q;
            zz[q] = j < nAct ? z[j] * invTemp : -INFINITY;
            mx = fmaxf(mx, zz[q]);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < 4; q++) { p[q] = (lane + 32 * q) < nAct ? expf(zz[q] - mx) : 0.f; sum += p[q]; }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        const int a = act[row];
        float ent = 0.f, lpa = 0.f, pca = 1.f;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = lane + 32 * q;
            p[q] *= inv;
            const float pc = fminf(fmaxf(p[q], kMinProb), 1.f);
            lp[q] = logf(pc);
            if (j < nAct) { ent -= lp[q] * pc; if (j == a) { lpa = lp[q]; pca = pc; } }
        }
        ent = warp_sum(ent);
        lpa = warp_sum(lpa);                                     // exactly one lane holds a
        pca = __shfl_sync(0xffffffffu, pca, a & 31);             // (of the lane's 4 slots only q = a / 32 matched)
        const float logRatio = lpa - oldLp[row];
        const float ratio = expf(logRatio);
        const float ad = adv[row];
        const float lo = 1.f - clipRange, hi = 1.f + clipRange;
        const float clipped = fminf(fmaxf(ratio, lo), hi);
        const bool inRange = ratio >= lo && ratio <= hi;
        const float gRatio = (inRange || ratio * ad < clipped * ad) ? ad : 0.f;
        const float gLp = -(ratioB * invN) * gRatio * ratio;
        const float gEnt = entCoef * ratioB * invN;
        float g[4], dot = 0.f;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = lane + 32 * q;
            g[q] = 0.f;
            if (j < nAct && p[q] >= kMinProb && p[q] <= 1.f) {
                g[q] = gEnt * (lp[q] + 1.f);
                if (j == a) g[q] += gLp / pca;
            }
            dot += g[q] * p[q];
        }
        dot = warp_sum(dot);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = lane + 32 * q;
            if (j < nActP) tile[r * ts + j] = j < nAct ? invTemp * p[q] * (g[q] - dot) : 0.f;
        }
        if (lane == 0) {
            sEnt += ent; sRatio += ratio;
            sKl += (expf(logRatio) - 1.f) - logRatio;
            sClip += fabsf(ratio - 1.f) > clipRange ? 1.f : 0.f;
        }
    }
    if (lane == 0) { red[warp][0] = sEnt; red[warp][1] = sKl; red[warp][2] = sRatio; red[warp][3] = sClip; }
    __syncthreads();
    const int rowsHere = (int)((n - row0) < 32 ? (n - row0) : 32);
    for (int i = threadIdx.x; i < rowsHere * nActP; i += 256) dZ[(size_t)row0 * nActP + i] = tile[(i / nActP) * ts + (i % nActP)];
    for (int j = warp; j < nActP; j += 8)
        if (lane < rowsHere) dZt[(size_t)j * ldT + row0 + lane] = tile[lane * ts + j];
    if (threadIdx.x < 4) {
        float s = 0.f;
        for (int w = 0; w < 8; w++) s += red[w][threadIdx.x];
        // accumulators hold per-minibatch MEANS summed over the minibatches (entropy, kl, ratio, clip fraction: acc 0, 1, 2, 4)
        const int slot = threadIdx.x == 3 ? 4 : threadIdx.x;
        atomicAdd(acc + slot, (double)s * (double)invN);
    }
}

// value loss = mean (v - target)^2 * ratioB (PPOLearner.cpp:200-205): dV = 2 (v - target) ratioB / n
__global__ void k_value_loss(const float* __restrict__ vals, int ldv, long n, long ldT, const float* __restrict__ tgt, float ratioB, float* __restrict__ dV,
                             float* __restrict__ dVt, double* __restrict__ acc) {
    __shared__ float red[8];
    pdl_enter();
    const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
    float sq = 0.f;
    if (row < n) {
        const float d = vals[(size_t)row * ldv] - tgt[row];
        const float gd = 2.f * d * ratioB / (float)n;
        float4 o = make_float4(gd, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(dV + (size_t)row * 4) = o;
        dVt[row] = gd;  // row 0 of [4, ldT]; rows 1..3 are the padding of the 1-wide output
        dVt[ldT + row] = 0.f; dVt[2 * ldT + row] = 0.f; dVt[3 * ldT + row] = 0.f;
        sq = d * d;
    }
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
        atomicAdd(acc + 3, (double)s * (double)ratioB / (double)n);
    }
}

// db[o] += sum over the rows of dY^T[o, :]
__global__ void __launch_bounds__(256) k_bias_grad(const float* __restrict__ dYt, long n, long ldT, float* __restrict__ db) {
    __shared__ float red[8];
    pdl_enter();
    const float* src = dYt + (size_t)blockIdx.x * ldT;
    float s = 0.f;
    const long n4 = n >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (long i = threadIdx.x; i < n4; i += 256) { float4 v = s4[i]; s += (v.x + v.y) + (v.z + v.w); }
    for (long i = (n4 << 2) + threadIdx.x; i < n; i += 256) s += src[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; w++) t += red[w];
        db[blockIdx.x] += t;
    }
}

// sum of squares of two segments [0, n0) and [n0, n0 + n1) of (a - b) (b may be NULL) -> out[0], out[1]
__global__ void __launch_bounds__(256) k_sumsq2(const float* __restrict__ a, const float* __restrict__ b, size_t n0, size_t n1, float scale,
                                                double* __restrict__ out) {
    __shared__ float red[2][8];
    pdl_enter();
    float s0 = 0.f, s1 = 0.f;
    const size_t total = n0 + n1;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        float v = (a[i] - (b ? b[i] : 0.f)) * scale;
        if (i < n0) s0 += v * v; else s1 += v * v;
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
        for (int w = 0; w < 8; w++) t += red[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, (double)t);
    }
}

// clip_grad_norm_(params, 0.5) per network + torch.optim.Adam (betas 0.9 / 0.999, eps 1e-8), every parameter of both networks
struct AdamNet { float lr, stepSize, bc2Sqrt; int32_t train; };
__global__ void k_adam(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n0, size_t n1, AdamNet a0,
                       AdamNet a1, float gradScale, float maxNorm, const double* __restrict__ normSq) {
    pdl_enter();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0 + n1) return;
    const int net = i < n0 ? 0 : 1;
    const AdamNet a = net == 0 ? a0 : a1;
    float grad = g[i] * gradScale;
    g[i] = 0.f;  // zero_grad for the next batch
    if (!a.train) return;
    const float norm = (float)sqrt(normSq[net]);
    const float coef = fminf(maxNorm / (norm + 1e-6f), 1.f);
    grad *= coef;
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const float mi = m[i] + (grad - m[i]) * (1.f - b1);      // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * b2 + (1.f - b2) * grad * grad;    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / a.bc2Sqrt + eps;
    p[i] -= a.stepSize * (mi / denom);
}

__global__ void k_transpose(const float* __restrict__ W, int rows, int cols, float* __restrict__ Wt) {  // Wt[c][r] = W[r][c]
    __shared__ float t[32][33];
    pdl_enter();
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) { int r = r0 + j, c = c0 + threadIdx.x; t[j][threadIdx.x] = (r < rows && c < cols) ? W[(size_t)r * cols + c] : 0.f; }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) { int c = c0 + j, r = r0 + threadIdx.x; if (r < rows && c < cols) Wt[(size_t)c * rows + r] = t[threadIdx.x][j]; }
}

struct PpoNet {
    int L = 0;
    int in[kMaxL] = {}, out[kMaxL] = {}, inP[kMaxL] = {}, outP[kMaxL] = {};
    size_t wOff[kMaxL] = {}, bOff[kMaxL] = {};  // float offsets into the flat vectors
    size_t begin = 0, count = 0;
    float* Wt[kMaxL] = {};
    float* Y[kMaxL] = {};   // forward outputs [n, outP]
    float* Yt[kMaxL] = {};  // their transposes [outP, ldT] (hidden layers)
    long step = 0;          // Adam step count
};

}  // namespace

struct rlg_ppo {
    rlg_ppo_cfg cfg;
    int device = 0;
    cudaStream_t own = nullptr;
    PpoNet net[2];
    size_t total = 0;
    float *params = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr, *before = nullptr;
    // experience ring (ExperienceBuffer)
    long cap = 0, cur = 0, head = 0;
    float* bStates = nullptr; int64_t* bActions = nullptr; float *bLogp = nullptr, *bTarget = nullptr, *bAdv = nullptr;
    // staging for rlg_ppo_submit_collector
    long stageRows = 0;
    float* sStates = nullptr; int64_t* sActions = nullptr; float *sLogp = nullptr, *sTarget = nullptr, *sAdv = nullptr;
    // minibatch workspace
    long mbRows = 0, ldT = 0;
    int obsP = 0, actP = 0, maxW = 0;
    float *X = nullptr, *Xt = nullptr; int32_t* act = nullptr; float *adv = nullptr, *oldLp = nullptr, *tgt = nullptr;
    float* dAct[2] = {nullptr, nullptr};   // ping-pong dY [n, maxW]
    float* dActT[2] = {nullptr, nullptr};  // ... and dY^T [maxW, ldT]
    // shuffle
    uint64_t *keysIn = nullptr, *keysOut = nullptr; int32_t *idxIn = nullptr, *perm = nullptr; void* cubTemp = nullptr; size_t cubBytes = 0;
    uint64_t shuffleCounter = 0;
    double* acc = nullptr;
    rlg_allreduce_hook hook = nullptr; void* hookUser = nullptr;
    uint64_t launches = 0;
    long updates = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

inline float* Wp(rlg_ppo* p, float* base, int n, int l) { return base + p->net[n].wOff[l]; }
inline float* Bp(rlg_ppo* p, float* base, int n, int l) { return base + p->net[n].bOff[l]; }

int refresh_transposes(rlg_ppo* p, cudaStream_t s) {
    for (int n = 0; n < 2; n++)
        for (int l = 1; l < p->net[n].L; l++) {
            const PpoNet& N = p->net[n];
            dim3 grid((N.inP[l] + 31) / 32, (N.outP[l] + 31) / 32);
            CKP(launch_pdl(k_transpose, grid, dim3(32, 8), 0, s, Wp(p, p->params, n, l), N.outP[l], N.inP[l], N.Wt[l]));
            p->launches++;
        }
    CKP(cudaGetLastError());
    return RLG_OK;
}

int forward_net(rlg_ppo* p, int n, long rows, cudaStream_t s) {
    PpoNet& N = p->net[n];
    const float* cur = p->X;
    int ld = p->obsP;
    for (int l = 0; l < N.L; l++) {
        const bool last = l == N.L - 1;
        CKR(rlg_gemm_tf32_fused((int)rows, N.outP[l], N.inP[l], cur, ld, Wp(p, p->params, n, l), N.inP[l], N.Y[l], N.outP[l], Bp(p, p->params, n, l),
                                last ? 0 : RLG_GEMM_RELU, 1, nullptr, 0, last ? nullptr : N.Yt[l], (int)p->ldT, s));
        p->launches++;
        cur = N.Y[l]; ld = N.outP[l];
    }
    return RLG_OK;
}

// dY (p->dAct[0], [rows, outP_last]) and dY^T (p->dActT[0]) hold the loss gradient of the net's output
int backward_net(rlg_ppo* p, int n, long rows, cudaStream_t s) {
    PpoNet& N = p->net[n];
    int cur = 0;
    const int K = (int)p->ldT;  // rows padded to a multiple of 4; the pad columns of every transposed operand are zero
    for (int l = N.L - 1; l >= 0; l--) {
        const float* dY = p->dAct[cur];
        const float* dYt = p->dActT[cur];
        const float* inT = l == 0 ? p->Xt : N.Yt[l - 1];
        int split = (int)(rows / 256);
        const int mTiles = (N.outP[l] + 127) / 128, nTiles = (N.inP[l] + 255) / 256;
        const int cap = 148 / (mTiles * nTiles > 0 ? mTiles * nTiles : 1);
        if (split > cap) split = cap;
        if (split < 1) split = 1;
        CKR(rlg_gemm_tf32_fused(N.outP[l], N.inP[l], K, dYt, K, inT, K, Wp(p, p->grads, n, l), N.inP[l], nullptr, RLG_GEMM_ATOMIC, split, nullptr, 0,
                                nullptr, 0, s));
        CKP(launch_pdl(k_bias_grad, dim3(N.outP[l]), dim3(256), 0, s, dYt, rows, p->ldT, Bp(p, p->grads, n, l)));
        p->launches += 2;
        if (l > 0) {
            CKR(rlg_gemm_tf32_fused((int)rows, N.inP[l], N.outP[l], dY, N.outP[l], N.Wt[l], N.outP[l], p->dAct[cur ^ 1], N.inP[l], nullptr, 0, 1,
                                    N.Y[l - 1], N.outP[l - 1], p->dActT[cur ^ 1], (int)p->ldT, s));
            p->launches++;
            cur ^= 1;
        }
    }
    CKP(cudaGetLastError());
    return RLG_OK;
}

void free_all(rlg_ppo* p) {
    cudaSetDevice(p->device);
    for (int n = 0; n < 2; n++)
        for (int l = 0; l < kMaxL; l++) { cudaFree(p->net[n].Wt[l]); cudaFree(p->net[n].Y[l]); cudaFree(p->net[n].Yt[l]); }
    cudaFree(p->params); cudaFree(p->grads); cudaFree(p->m); cudaFree(p->v); cudaFree(p->before);
    cudaFree(p->bStates); cudaFree(p->bActions); cudaFree(p->bLogp); cudaFree(p->bTarget); cudaFree(p->bAdv);
    cudaFree(p->sStates); cudaFree(p->sActions); cudaFree(p->sLogp); cudaFree(p->sTarget); cudaFree(p->sAdv);
    cudaFree(p->X); cudaFree(p->Xt); cudaFree(p->act); cudaFree(p->adv); cudaFree(p->oldLp); cudaFree(p->tgt);
    for (int i = 0; i < 2; i++) { cudaFree(p->dAct[i]); cudaFree(p->dActT[i]); }
    cudaFree(p->keysIn); cudaFree(p->keysOut); cudaFree(p->idxIn); cudaFree(p->perm); cudaFree(p->cubTemp); cudaFree(p->acc);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->own) cudaStreamDestroy(p->own);
}

}  // namespace

extern "C" {

int rlg_ppo_destroy(rlg_ppo* p) {
    if (!p) return RLG_OK;
    free_all(p);
    delete p;
    return RLG_OK;
}

int rlg_ppo_create(const rlg_ppo_cfg* cfg, rlg_ppo** out) {
    if (!cfg || !out) return failp(RLG_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->num_hidden < 1 || cfg->num_hidden > RLG_MAX_HIDDEN_LAYERS) return failp(RLG_ERR_INVALID, "num_hidden must be 1..4");
    if (cfg->obs_size < 1 || cfg->num_actions < 1 || cfg->num_actions > 128) return failp(RLG_ERR_INVALID, "bad obs_size / num_actions (<= 128)");
    if (cfg->batch_size < 1 || cfg->exp_buffer_size < 1) return failp(RLG_ERR_INVALID, "batch_size and exp_buffer_size must be positive");
    long mbs = cfg->mini_batch_size == 0 ? cfg->batch_size : cfg->mini_batch_size;  // PPOLearner.cpp:19-20
    if (mbs < 1 || cfg->batch_size % mbs != 0) return failp(RLG_ERR_INVALID, "PPOLearner: batchSize must be a multiple of miniBatchSize");  // :22-23
    if (!(cfg->temperature > 0.f)) return failp(RLG_ERR_INVALID, "temperature must be > 0");
    for (int i = 0; i < cfg->num_hidden; i++)
        for (int h : {cfg->policy_hidden[i], cfg->critic_hidden[i]})
            if (h < 4 || (h & 3)) return failp(RLG_ERR_INVALID, "hidden layer sizes must be multiples of 4");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return failp(RLG_ERR_CUDA, "no CUDA device available: the PPO learner has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return failp(RLG_ERR_INVALID, "bad device ordinal");
    rlg_ppo* p = new (std::nothrow) rlg_ppo();
    if (!p) return failp(RLG_ERR_INVALID, "out of host memory");
    p->cfg = *cfg;
    p->cfg.mini_batch_size = mbs;
    if (p->cfg.world < 1) p->cfg.world = 1;
    p->device = cfg->device;
    p->obsP = pad4(cfg->obs_size);
    p->actP = pad4(cfg->num_actions);
    p->mbRows = mbs;
    p->ldT = (mbs + 3) & ~3L;
    p->cap = cfg->exp_buffer_size;
    size_t off = 0;
    int maxW = p->actP > 4 ? p->actP : 4;
    for (int n = 0; n < 2; n++) {
        PpoNet& N = p->net[n];
        N.L = cfg->num_hidden + 1;
        N.begin = off;
        int in = cfg->obs_size, inP = p->obsP;
        for (int l = 0; l < N.L; l++) {
            const int o = l < cfg->num_hidden ? (n == 0 ? cfg->policy_hidden[l] : cfg->critic_hidden[l]) : (n == 0 ? cfg->num_actions : 1);
            N.in[l] = in; N.inP[l] = inP; N.out[l] = o; N.outP[l] = pad4(o);
            N.wOff[l] = off; off += (size_t)N.outP[l] * N.inP[l];
            N.bOff[l] = off; off += (size_t)N.outP[l];
            if (N.outP[l] > maxW) maxW = N.outP[l];
            in = o; inP = N.outP[l];
        }
        N.count = off - N.begin;
    }
    p->total = off;
    p->maxW = maxW;
#define CKX(expr)                                                                                                                            \
    do {                                                                                                                                     \
        cudaError_t _e = (expr);                                                                                                             \
        if (_e != cudaSuccess) { std::string msg = std::string(#expr) + ": " + cudaGetErrorString(_e); rlg_ppo_destroy(p); return failp(RLG_ERR_CUDA, msg); } \
    } while (0)
    CKX(cudaSetDevice(p->device));
    CKX(cudaStreamCreateWithFlags(&p->own, cudaStreamNonBlocking));
    CKX(cudaEventCreate(&p->ev0));
    CKX(cudaEventCreate(&p->ev1));
    for (float** a : {&p->params, &p->grads, &p->m, &p->v, &p->before}) {
        CKX(cudaMalloc(a, p->total * 4));
        CKX(cudaMemset(*a, 0, p->total * 4));
    }
    const size_t rows = (size_t)mbs, ldT = (size_t)p->ldT;
    for (int n = 0; n < 2; n++) {
        PpoNet& N = p->net[n];
        for (int l = 0; l < N.L; l++) {
            CKX(cudaMalloc(&N.Y[l], rows * N.outP[l] * 4));
            if (l < N.L - 1) { CKX(cudaMalloc(&N.Yt[l], ldT * N.outP[l] * 4)); CKX(cudaMemset(N.Yt[l], 0, ldT * N.outP[l] * 4)); }
            if (l > 0) CKX(cudaMalloc(&N.Wt[l], (size_t)N.outP[l] * N.inP[l] * 4));
        }
    }
    CKX(cudaMalloc(&p->X, rows * p->obsP * 4));
    CKX(cudaMalloc(&p->Xt, ldT * p->obsP * 4));
    CKX(cudaMemset(p->Xt, 0, ldT * p->obsP * 4));
    CKX(cudaMalloc(&p->act, rows * 4)); CKX(cudaMalloc(&p->adv, rows * 4)); CKX(cudaMalloc(&p->oldLp, rows * 4)); CKX(cudaMalloc(&p->tgt, rows * 4));
    for (int i = 0; i < 2; i++) {
        CKX(cudaMalloc(&p->dAct[i], rows * maxW * 4));
        CKX(cudaMalloc(&p->dActT[i], ldT * maxW * 4));
        CKX(cudaMemset(p->dActT[i], 0, ldT * maxW * 4));
    }
    const size_t cap = (size_t)p->cap;
    CKX(cudaMalloc(&p->bStates, cap * cfg->obs_size * 4));
    CKX(cudaMalloc(&p->bActions, cap * 8)); CKX(cudaMalloc(&p->bLogp, cap * 4)); CKX(cudaMalloc(&p->bTarget, cap * 4)); CKX(cudaMalloc(&p->bAdv, cap * 4));
    CKX(cudaMalloc(&p->keysIn, cap * 8)); CKX(cudaMalloc(&p->keysOut, cap * 8)); CKX(cudaMalloc(&p->idxIn, cap * 4)); CKX(cudaMalloc(&p->perm, cap * 4));
    CKX(cub::DeviceRadixSort::SortPairs(nullptr, p->cubBytes, p->keysIn, p->keysOut, p->idxIn, p->perm, (int)cap));
    CKX(cudaMalloc(&p->cubTemp, p->cubBytes));
    CKX(cudaMalloc(&p->acc, kAccWords * 8));
    CKX(cudaMemset(p->acc, 0, kAccWords * 8));
#undef CKX
    *out = p;
    return RLG_OK;
}

void* rlg_ppo_stream(rlg_ppo* p) { return p ? (void*)p->own : nullptr; }
uint64_t rlg_ppo_launch_count(const rlg_ppo* p) { return p ? p->launches : 0; }
int64_t rlg_ppo_buffer_size(const rlg_ppo* p) { return p ? p->cur : 0; }
int64_t rlg_ppo_model_updates(const rlg_ppo* p) { return p ? p->updates : 0; }

int rlg_ppo_set_allreduce_hook(rlg_ppo* p, rlg_allreduce_hook hook, void* user, int world) {
    if (!p || world < 1) return failp(RLG_ERR_INVALID, "bad argument");
    p->hook = hook; p->hookUser = user; p->cfg.world = world;
    return RLG_OK;
}

int rlg_ppo_set_lr(rlg_ppo* p, float policy_lr, float critic_lr) {  // PPOLearner::UpdateLearningRates (PPOLearner.cpp:504-517)
    if (!p) return failp(RLG_ERR_INVALID, "null learner");
    p->cfg.policy_lr = policy_lr; p->cfg.critic_lr = critic_lr;
    return RLG_OK;
}

int rlg_ppo_flat(rlg_ppo* p, int which, float** dev, int64_t* count, int64_t* policy_count) {
    if (!p || which < 0 || which > 3) return failp(RLG_ERR_INVALID, "bad argument");
    float* a[4] = {p->params, p->grads, p->m, p->v};
    if (dev) *dev = a[which];
    if (count) *count = (int64_t)p->total;
    if (policy_count) *policy_count = (int64_t)p->net[0].count;
    return RLG_OK;
}

static int layer_io(rlg_ppo* p, int which, int net, int layer, float* W, float* b, int out_dim, int in_dim, bool set) {
    if (!p || (!W && !b)) return failp(RLG_ERR_INVALID, "null argument");
    if (which < 0 || which > 3 || net < 0 || net > 1 || layer < 0 || layer >= p->net[net].L) return failp(RLG_ERR_INVALID, "bad which / net / layer index");
    const PpoNet& N = p->net[net];
    if (out_dim != N.out[layer] || in_dim != N.in[layer]) return failp(RLG_ERR_INVALID, "layer shape does not match the learner configuration");
    CKP(cudaSetDevice(p->device));
    float* bases[4] = {p->params, p->grads, p->m, p->v};
    float* dW = bases[which] + N.wOff[layer];
    float* dB = bases[which] + N.bOff[layer];
    CKP(cudaStreamSynchronize(p->own));
    if (set) {
        if (W) CKP(cudaMemcpy2D(dW, (size_t)N.inP[layer] * 4, W, (size_t)in_dim * 4, (size_t)in_dim * 4, out_dim, cudaMemcpyHostToDevice));
        if (b) CKP(cudaMemcpy(dB, b, (size_t)out_dim * 4, cudaMemcpyHostToDevice));
        if (which == 0 && layer > 0 && W) {
            dim3 grid((N.inP[layer] + 31) / 32, (N.outP[layer] + 31) / 32);
            CKP(launch_pdl(k_transpose, grid, dim3(32, 8), 0, p->own, dW, N.outP[layer], N.inP[layer], N.Wt[layer]));
            CKP(cudaStreamSynchronize(p->own));
        }
    } else {
        if (W) CKP(cudaMemcpy2D(W, (size_t)in_dim * 4, dW, (size_t)N.inP[layer] * 4, (size_t)in_dim * 4, out_dim, cudaMemcpyDeviceToHost));
        if (b) CKP(cudaMemcpy(b, dB, (size_t)out_dim * 4, cudaMemcpyDeviceToHost));
    }
    return RLG_OK;
}
int rlg_ppo_set_layer(rlg_ppo* p, int which, int net, int layer, const float* W_host, const float* b_host, int out_dim, int in_dim) {
    return layer_io(p, which, net, layer, const_cast<float*>(W_host), const_cast<float*>(b_host), out_dim, in_dim, true);
}
int rlg_ppo_get_layer(rlg_ppo* p, int which, int net, int layer, float* W_host, float* b_host, int out_dim, int in_dim) {
    return layer_io(p, which, net, layer, W_host, b_host, out_dim, in_dim, false);
}
int rlg_ppo_adam_steps(rlg_ppo* p, int64_t* policy_steps, int64_t* critic_steps, int set) {
    if (!p || !policy_steps || !critic_steps) return failp(RLG_ERR_INVALID, "null argument");
    if (set) { p->net[0].step = (long)*policy_steps; p->net[1].step = (long)*critic_steps; }
    else { *policy_steps = p->net[0].step; *critic_steps = p->net[1].step; }
    return RLG_OK;
}

// torch::nn::Linear's default initialisation (kaiming_uniform_(a = sqrt 5) on the weight, U(-1/sqrt(fan_in), 1/sqrt(fan_in)) on the
// bias: both are U(-1/sqrt(in), 1/sqrt(in))), which DiscretePolicy / ValueEstimator keep (DiscretePolicy.cpp:13-27).  The random
// stream is a host mt19937_64 — the distribution equals torch's, the numbers do not.
int rlg_ppo_init_weights(rlg_ppo* p, uint64_t seed) {
    if (!p) return failp(RLG_ERR_INVALID, "null learner");
    std::mt19937_64 gen(seed);
    for (int n = 0; n < 2; n++)
        for (int l = 0; l < p->net[n].L; l++) {
            const PpoNet& N = p->net[n];
            const float bound = 1.f / std::sqrt((float)N.in[l]);
            std::uniform_real_distribution<float> U(-bound, bound);
            std::vector<float> W((size_t)N.out[l] * N.in[l]), b((size_t)N.out[l]);
            for (auto& x : W) x = U(gen);
            for (auto& x : b) x = U(gen);
            CKR(rlg_ppo_set_layer(p, 0, n, l, W.data(), b.data(), N.out[l], N.in[l]));
        }
    return RLG_OK;
}

// ExperienceBuffer::SubmitExperience (ExperienceBuffer.cpp:12-70): FIFO of exp_buffer_size rows; the oldest rows fall out.
int rlg_ppo_submit(rlg_ppo* p, const float* states, const int64_t* actions, const float* log_probs, const float* value_targets,
                   const float* advantages, int64_t n, void* stream) {
    if (!p || !states || !actions || !log_probs || !value_targets || !advantages || n < 0) return failp(RLG_ERR_INVALID, "bad argument");
    if (n == 0) return RLG_OK;
    CKP(cudaSetDevice(p->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : p->own;
    const long cap = p->cap, obs = p->cfg.obs_size;
    long skip = 0;
    if (n > cap) { skip = n - cap; n = cap; }  // only the newest cap rows survive
    const long overflow = p->cur + n - cap > 0 ? p->cur + n - cap : 0;
    p->head = (p->head + overflow) % cap;
    p->cur -= overflow;
    long dst = (p->head + p->cur) % cap;
    long done = 0;
    while (done < n) {
        const long chunk = (n - done) < (cap - dst) ? (n - done) : (cap - dst);
        const long src = skip + done;
        CKP(cudaMemcpyAsync(p->bStates + (size_t)dst * obs, states + (size_t)src * obs, (size_t)chunk * obs * 4, cudaMemcpyDeviceToDevice, s));
        CKP(cudaMemcpyAsync(p->bActions + dst, actions + src, (size_t)chunk * 8, cudaMemcpyDeviceToDevice, s));
        CKP(cudaMemcpyAsync(p->bLogp + dst, log_probs + src, (size_t)chunk * 4, cudaMemcpyDeviceToDevice, s));
        CKP(cudaMemcpyAsync(p->bTarget + dst, value_targets + src, (size_t)chunk * 4, cudaMemcpyDeviceToDevice, s));
        CKP(cudaMemcpyAsync(p->bAdv + dst, advantages + src, (size_t)chunk * 4, cudaMemcpyDeviceToDevice, s));
        done += chunk;
        dst = (dst + chunk) % cap;
    }
    p->cur += n;
    return RLG_OK;
}

// Learner::AddNewExperience's hand-over (Learner.cpp:684-701): the collector's last collect, in the reference's row order, into the FIFO
int rlg_ppo_submit_collector(rlg_ppo* p, rlg_collector* c, void* stream) {
    if (!p || !c) return failp(RLG_ERR_INVALID, "null argument");
    rlg_traj_view v;
    CKR(rlg_collector_view(c, &v));
    if (v.obs_size != p->cfg.obs_size) return failp(RLG_ERR_INVALID, "collector obs size does not match the learner");
    const long rows = (long)v.T * v.N;
    CKP(cudaSetDevice(p->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : p->own;
    if (rows > p->stageRows) {
        CKP(cudaStreamSynchronize(s));
        cudaFree(p->sStates); cudaFree(p->sActions); cudaFree(p->sLogp); cudaFree(p->sTarget); cudaFree(p->sAdv);
        p->sStates = nullptr; p->sActions = nullptr; p->sLogp = p->sTarget = p->sAdv = nullptr; p->stageRows = 0;
        CKP(cudaMalloc(&p->sStates, (size_t)rows * v.obs_size * 4));
        CKP(cudaMalloc(&p->sActions, (size_t)rows * 8)); CKP(cudaMalloc(&p->sLogp, (size_t)rows * 4));
        CKP(cudaMalloc(&p->sTarget, (size_t)rows * 4)); CKP(cudaMalloc(&p->sAdv, (size_t)rows * 4));
        p->stageRows = rows;
    }
    CKR(rlg_collector_export(c, p->sStates, p->sActions, p->sLogp, nullptr, nullptr, nullptr, nullptr, p->sTarget, p->sAdv, s));
    return rlg_ppo_submit(p, p->sStates, p->sActions, p->sLogp, p->sTarget, p->sAdv, rows, s);
}

// Logical (FIFO-order) read-back of the buffer for tests / checkpoint tools: any pointer may be NULL.
int rlg_ppo_buffer_read(rlg_ppo* p, float* states_host, int64_t* actions_host, float* log_probs_host, float* value_targets_host, float* advantages_host) {
    if (!p) return failp(RLG_ERR_INVALID, "null learner");
    CKP(cudaSetDevice(p->device));
    CKP(cudaDeviceSynchronize());
    const long cap = p->cap, obs = p->cfg.obs_size;
    long done = 0, src = p->head;
    while (done < p->cur) {
        const long chunk = (p->cur - done) < (cap - src) ? (p->cur - done) : (cap - src);
        if (states_host) CKP(cudaMemcpy(states_host + (size_t)done * obs, p->bStates + (size_t)src * obs, (size_t)chunk * obs * 4, cudaMemcpyDeviceToHost));
        if (actions_host) CKP(cudaMemcpy(actions_host + done, p->bActions + src, (size_t)chunk * 8, cudaMemcpyDeviceToHost));
        if (log_probs_host) CKP(cudaMemcpy(log_probs_host + done, p->bLogp + src, (size_t)chunk * 4, cudaMemcpyDeviceToHost));
        if (value_targets_host) CKP(cudaMemcpy(value_targets_host + done, p->bTarget + src, (size_t)chunk * 4, cudaMemcpyDeviceToHost));
        if (advantages_host) CKP(cudaMemcpy(advantages_host + done, p->bAdv + src, (size_t)chunk * 4, cudaMemcpyDeviceToHost));
        done += chunk;
        src = (src + chunk) % cap;
    }
    return RLG_OK;
}

// The permutation of the NEXT epoch (tests replay the learner's batches with it): perm_host [buffer_size] int32 of logical rows.
int rlg_ppo_peek_shuffle(rlg_ppo* p, int32_t* perm_host, uint64_t counter) {
    if (!p || !perm_host) return failp(RLG_ERR_INVALID, "null argument");
    if (p->cur < 1) return RLG_OK;
    CKP(cudaSetDevice(p->device));
    const long n = p->cur;
    CKP(launch_pdl(k_shuffle_keys, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, p->own, p->keysIn, p->idxIn, n, p->cfg.seed, counter));
    CKP(cub::DeviceRadixSort::SortPairs(p->cubTemp, p->cubBytes, p->keysIn, p->keysOut, p->idxIn, p->perm, (int)n, 0, 64, p->own));
    CKP(cudaMemcpyAsync(perm_host, p->perm, (size_t)n * 4, cudaMemcpyDeviceToHost, p->own));
    CKP(cudaStreamSynchronize(p->own));
    return RLG_OK;
}
uint64_t rlg_ppo_shuffle_counter(const rlg_ppo* p) { return p ? p->shuffleCounter : 0; }

// PPOLearner::Learn (PPOLearner.cpp:67-349)
int rlg_ppo_learn(rlg_ppo* p, rlg_ppo_report* report, void* stream) {
    if (!p) return failp(RLG_ERR_INVALID, "null learner");
    CKP(cudaSetDevice(p->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : p->own;
    const rlg_ppo_cfg& cfg = p->cfg;
    const bool trainPolicy = cfg.policy_lr != 0.f, trainCritic = cfg.critic_lr != 0.f;  // PPOLearner.cpp:84-85
    const long batch = cfg.batch_size, mbs = cfg.mini_batch_size;
    const float ratioB = (float)mbs / (float)batch;  // batchSizeRatio, PPOLearner.cpp:131
    const size_t n0 = p->net[0].count, n1 = p->net[1].count;
    long nBatches = 0, nMini = 0;
    CKP(cudaEventRecord(p->ev0, s));
    CKP(cudaMemcpyAsync(p->before, p->params, p->total * 4, cudaMemcpyDeviceToDevice, s));
    CKP(cudaMemsetAsync(p->acc, 0, kAccWords * 8, s));
    CKP(cudaMemsetAsync(p->grads, 0, p->total * 4, s));
    const int gatherSmem = 32 * (p->obsP + 1) * 4, lossSmem = 32 * (p->actP + 1) * 4;
    for (int epoch = 0; epoch < cfg.epochs; epoch++) {
        const long n = p->cur;
        if (n < batch) break;  // full batches only (ExperienceBuffer.cpp:114)
        CKP(launch_pdl(k_shuffle_keys, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, p->keysIn, p->idxIn, n, cfg.seed, p->shuffleCounter++));
        CKP(cub::DeviceRadixSort::SortPairs(p->cubTemp, p->cubBytes, p->keysIn, p->keysOut, p->idxIn, p->perm, (int)n, 0, 64, s));
        p->launches += 2;
        for (long b0 = 0; b0 + batch <= n; b0 += batch) {
            for (long m0 = 0; m0 < batch; m0 += mbs) {
                CKP(launch_pdl(k_gather_rows, dim3((unsigned)((mbs + 31) / 32)), dim3(256), gatherSmem, s, p->perm + b0 + m0, mbs, p->ldT, p->head, p->cap, cfg.obs_size, p->obsP,
                                                                                 p->bStates, p->bActions, p->bLogp, p->bTarget, p->bAdv, p->X, p->Xt, p->act,
                                                                                 p->oldLp, p->tgt, p->adv));
                p->launches++;
                if (trainCritic) {
                    CKR(forward_net(p, 1, mbs, s));
                    const PpoNet& C = p->net[1];
                    CKP(launch_pdl(k_value_loss, dim3((unsigned)((mbs + 255) / 256)), dim3(256), 0, s, C.Y[C.L - 1], C.outP[C.L - 1], mbs, p->ldT, p->tgt, ratioB, p->dAct[0], p->dActT[0],
                                                                              p->acc));
                    p->launches++;
                    CKR(backward_net(p, 1, mbs, s));
                }
                if (trainPolicy) {
                    CKR(forward_net(p, 0, mbs, s));
                    const PpoNet& P = p->net[0];
                    CKP(launch_pdl(k_policy_loss, dim3((unsigned)((mbs + 31) / 32)), dim3(256), lossSmem, s, P.Y[P.L - 1], P.outP[P.L - 1], cfg.num_actions, p->actP, mbs, p->ldT, p->act,
                                                                                   p->adv, p->oldLp, 1.f / cfg.temperature, cfg.clip_range, cfg.ent_coef, ratioB,
                                                                                   p->dAct[0], p->dActT[0], p->acc));
                    p->launches++;
                    CKR(backward_net(p, 0, mbs, s));
                }
                nMini++;
            }
            // one collective for every gradient of both networks (SUM; the mean is taken below)
            if (p->hook && cfg.world > 1) p->hook(p->hookUser, p->grads, (int64_t)p->total, (void*)s);
            const float gradScale = 1.f / (float)cfg.world;
            CKP(cudaMemsetAsync(p->acc + 5, 0, 16, s));
            CKP(launch_pdl(k_sumsq2, dim3(148), dim3(256), 0, s, p->grads, nullptr, n0, n1, gradScale, p->acc + 5));
            AdamNet a[2];
            for (int net = 0; net < 2; net++) {
                const bool train = net == 0 ? trainPolicy : trainCritic;
                if (train) p->net[net].step++;
                const double t = (double)(p->net[net].step > 0 ? p->net[net].step : 1);
                const double lr = net == 0 ? cfg.policy_lr : cfg.critic_lr;
                a[net].lr = (float)lr;
                a[net].stepSize = (float)(lr / (1.0 - std::pow(0.9, t)));
                a[net].bc2Sqrt = (float)std::sqrt(1.0 - std::pow(0.999, t));
                a[net].train = train ? 1 : 0;
            }
            CKP(launch_pdl(k_adam, dim3((unsigned)((p->total + 255) / 256)), dim3(256), 0, s, p->params, p->grads, p->m, p->v, n0, n1, a[0], a[1], gradScale, 0.5f, p->acc + 5));
            p->launches += 2;
            CKR(refresh_transposes(p, s));
            nBatches++;
        }
    }
    CKP(cudaMemsetAsync(p->acc + 7, 0, 16, s));
    CKP(launch_pdl(k_sumsq2, dim3(148), dim3(256), 0, s, p->params, p->before, n0, n1, 1.f, p->acc + 7));
    p->launches++;
    CKP(cudaEventRecord(p->ev1, s));
    CKP(cudaGetLastError());
    p->updates += nBatches;
    if (report) {
        double h[kAccWords];
        CKP(cudaMemcpyAsync(h, p->acc, sizeof(h), cudaMemcpyDeviceToHost, s));
        CKP(cudaStreamSynchronize(s));  // the one synchronisation of a learn call
        float ms = 0.f;
        CKP(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
        const double nm = nMini > 0 ? (double)nMini : 1.0;
        memset(report, 0, sizeof(*report));
        report->batches = nBatches; report->minibatches = nMini;
        report->entropy = h[0] / nm; report->kl = h[1] / nm; report->ratio = h[2] / nm; report->value_loss = h[3] / nm;
        report->clip_fraction = trainPolicy ? h[4] / nm : 0.0;
        report->policy_update_magnitude = std::sqrt(h[7]); report->critic_update_magnitude = std::sqrt(h[8]);
        report->device_ms = ms;
        report->cumulative_model_updates = p->updates;
    }
    return RLG_OK;
}

// ThreadAgentManager::SetNewPolicy / the critic hand-over: device-to-device repack of the current weights into the collector's
// inference layout (no host round trip).
int rlg_ppo_push_weights(rlg_ppo* p, rlg_collector* c, void* stream) {
    if (!p || !c) return failp(RLG_ERR_INVALID, "null argument");
    cudaStream_t s = stream ? (cudaStream_t)stream : p->own;
    for (int n = 0; n < 2; n++)
        for (int l = 0; l < p->net[n].L; l++) {
            const PpoNet& N = p->net[n];
            CKR(rlg_collector_set_layer_device(c, n, l, p->params + N.wOff[l], N.inP[l], p->params + N.bOff[l], N.out[l], N.in[l], (void*)s));
        }
    return RLG_OK;
}

}  // extern "C"
