// rl_math.h — small fp32 vector/matrix/quaternion kit shared by every kernel.
//
// Conventions follow the reference so tolerances stay tight:
//  * M3 is row-major like btMatrix3x3; a body's basis has forward/right/up as COLUMNS
//    (reference RocketSim/src/Math/MathTypes/MathTypes.h:171-178).
//  * dot() associates as (x*x' + y*y') + z*z' (reference bullet LinearMath/btVector3.h:230-247).
//  * The physics may be compiled with FMA contraction and approximate division / square root (parity there is
//    tolerance based); everything that must be BIT-EXACT against the FMA-free x86 reference build (the gym layer:
//    obs, rewards, event tracker) goes through the s_* helpers below, which are IEEE round-to-nearest single
//    operations that the compiler never fuses or approximates, whatever the flags.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RL_HD __host__ __device__
#define RL_HDI __host__ __device__ __forceinline__
#define RL_NOINLINE __noinline__
#else
#define RL_HD
#define RL_HDI inline
#define RL_NOINLINE __attribute__((noinline))
#endif

// the four per-wheel loops: rolled by default (the role kernel is instruction-cache bound); -DRL_UNROLL_WHEELS trades code
// size for instruction-level parallelism across the wheels (A/B: profiles/)
#ifdef RL_UNROLL_WHEELS
#define RL_WHEEL_LOOP _Pragma("unroll")
#else
#define RL_WHEEL_LOOP _Pragma("unroll 1")
#endif

// Diagnostic builds (-DRLG_PHASE_TIMING, tools/phase_prof.py): per-warp cycle counters of the sub-phases of a tick.
// RL_PT(i) adds the cycles since the warp's previous mark to slot i; no-ops in the product build.
#if defined(RLG_PHASE_TIMING) && defined(__CUDACC__)
static __device__ uint32_t* g_rl_pt;  // [block][warp][32]; slot 31 = last time stamp
#endif
#if defined(RLG_PHASE_TIMING) && defined(__CUDA_ARCH__)
__device__ __forceinline__ void rl_pt_mark(int i) {
    if ((int)(threadIdx.x & 31) != __ffs(__activemask()) - 1 || !g_rl_pt) return;
    uint32_t* b = g_rl_pt + ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    uint32_t now = (uint32_t)clock64();
    if (i >= 0) b[i] += now - b[31];
    b[31] = now;
}
#define RL_PT(i) rl_pt_mark(i)
#else
#define RL_PT(i) do {} while (0)
#endif

namespace rl {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kHalfPi = 1.57079632679489661923f;
constexpr float kEps = 1.1920928955078125e-7f;  // FLT_EPSILON == SIMD_EPSILON

// libm calls go through out-of-line wrappers: the device versions carry long argument-reduction slow paths and the role
// kernel is instruction-cache bound, so each must exist ONCE in the image (same results as the inlined calls).
RL_HD RL_NOINLINE float rl_sin(float x);
RL_HD RL_NOINLINE float rl_cos(float x);
RL_HD RL_NOINLINE float rl_atan2(float y, float x);
RL_HD RL_NOINLINE float rl_asin(float x);

// strict single operations (never contracted into FMA, never approximated)
#if defined(__CUDA_ARCH__)
RL_HDI float s_mul(float a, float b) { return __fmul_rn(a, b); }
RL_HDI float s_add(float a, float b) { return __fadd_rn(a, b); }
RL_HDI float s_sub(float a, float b) { return __fsub_rn(a, b); }
RL_HDI float s_div(float a, float b) { return __fdiv_rn(a, b); }
RL_HDI float s_sqrt(float a) { return __fsqrt_rn(a); }
#else
RL_HDI float s_mul(float a, float b) { return a * b; }
RL_HDI float s_add(float a, float b) { return a + b; }
RL_HDI float s_sub(float a, float b) { return a - b; }
RL_HDI float s_div(float a, float b) { return a / b; }
RL_HDI float s_sqrt(float a) { return sqrtf(a); }
#endif

// Tag for constructors that leave the storage uninitialised: per-thread arrays of solver rows / contacts / simplex
// vertices live in local memory, and zero-filling them on every call was 25 % of the role kernel's local-memory traffic
// (profiles/r01e_*).  Every member is written before it is read.
struct NoInit {};

struct V3 {
    float x, y, z;
    RL_HDI V3() : x(0), y(0), z(0) {}
    RL_HDI explicit V3(NoInit) {}
    RL_HDI V3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    RL_HDI float& operator[](int i) { return (&x)[i]; }
    RL_HDI float operator[](int i) const { return (&x)[i]; }
};

RL_HDI V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
RL_HDI V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
RL_HDI V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
RL_HDI V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
RL_HDI V3 operator*(float s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
RL_HDI V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
RL_HDI V3 operator/(V3 a, float s) { float r = 1.0f / s; return V3(a.x * r, a.y * r, a.z * r); }
RL_HDI V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
RL_HDI V3& operator-=(V3& a, V3 b) { a = a - b; return a; }
RL_HDI V3& operator*=(V3& a, float s) { a = a * s; return a; }
RL_HDI float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
RL_HDI V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
RL_HDI float len2(V3 a) { return dot(a, a); }
RL_HDI float len(V3 a) { return sqrtf(dot(a, a)); }
RL_HDI bool is_zero(V3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
RL_HDI V3 normalized(V3 a) { return a / len(a); }
// btVector3::safeNormalize (btVector3.h:287-300)
RL_HDI V3 safe_normalized(V3 a) {
    float l2 = len2(a);
    if (l2 >= kEps * kEps) return a / sqrtf(l2);
    return V3(1, 0, 0);
}
RL_HDI float fminf_(float a, float b) { return a < b ? a : b; }
RL_HDI float fmaxf_(float a, float b) { return a > b ? a : b; }
RL_HDI float clampf(float v, float lo, float hi) { return fminf_(fmaxf_(v, lo), hi); }
RL_HDI V3 vmin(V3 a, V3 b) { return V3(fminf_(a.x, b.x), fminf_(a.y, b.y), fminf_(a.z, b.z)); }
RL_HDI V3 vmax(V3 a, V3 b) { return V3(fmaxf_(a.x, b.x), fmaxf_(a.y, b.y), fmaxf_(a.z, b.z)); }
RL_HDI V3 vabs(V3 a) { return V3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
RL_HDI int sgn(float v) { return (v > 0.f) - (v < 0.f); }

struct M3 {
    V3 r[3];  // rows
    RL_HDI M3() {}
    RL_HDI explicit M3(NoInit) : r{V3(NoInit()), V3(NoInit()), V3(NoInit())} {}
    RL_HDI M3(V3 r0, V3 r1, V3 r2) { r[0] = r0; r[1] = r1; r[2] = r2; }
    RL_HDI V3 col(int i) const { return V3(r[0][i], r[1][i], r[2][i]); }
    RL_HDI static M3 identity() { return M3(V3(1, 0, 0), V3(0, 1, 0), V3(0, 0, 1)); }
    RL_HDI static M3 from_cols(V3 c0, V3 c1, V3 c2) {
        return M3(V3(c0.x, c1.x, c2.x), V3(c0.y, c1.y, c2.y), V3(c0.z, c1.z, c2.z));
    }
};
RL_HDI V3 operator*(const M3& m, V3 v) { return V3(dot(m.r[0], v), dot(m.r[1], v), dot(m.r[2], v)); }
// v^T * M  (btVector3 * btMatrix3x3)
RL_HDI V3 tmul(V3 v, const M3& m) { return V3(dot(m.col(0), v), dot(m.col(1), v), dot(m.col(2), v)); }
RL_HDI M3 transpose(const M3& m) { return M3(m.col(0), m.col(1), m.col(2)); }
RL_HDI M3 operator*(const M3& a, const M3& b) {
    M3 o;
    for (int i = 0; i < 3; i++) o.r[i] = V3(dot(a.r[i], b.col(0)), dot(a.r[i], b.col(1)), dot(a.r[i], b.col(2)));
    return o;
}
// btMatrix3x3::scaled: column i scaled by s[i]
RL_HDI M3 scaled(const M3& m, V3 s) { return M3(m.r[0] * s, m.r[1] * s, m.r[2] * s); }

// R * diag(d) * R^T  (btRigidBody::updateInertiaTensor, btRigidBody.cpp:258-261)
RL_HDI M3 world_inertia(const M3& rot, V3 d) { return scaled(rot, d) * transpose(rot); }

struct Quat {
    float x, y, z, w;
    RL_HDI Quat() : x(0), y(0), z(0), w(1) {}
    RL_HDI Quat(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
};
RL_HDI Quat operator*(Quat q1, Quat q2) {
    return Quat(q1.w * q2.x + q1.x * q2.w + q1.y * q2.z - q1.z * q2.y,
                q1.w * q2.y + q1.y * q2.w + q1.z * q2.x - q1.x * q2.z,
                q1.w * q2.z + q1.z * q2.w + q1.x * q2.y - q1.y * q2.x,
                q1.w * q2.w - q1.x * q2.x - q1.y * q2.y - q1.z * q2.z);
}
// btMatrix3x3::getRotation (LinearMath/btMatrix3x3.h, scalar path)
RL_HDI Quat mat_to_quat(const M3& m) {
    float trace = m.r[0].x + m.r[1].y + m.r[2].z;
    float t[4];
    if (trace > 0.f) {
        float s = sqrtf(trace + 1.0f);
        t[3] = s * 0.5f;
        s = 0.5f / s;
        t[0] = (m.r[2].y - m.r[1].z) * s;
        t[1] = (m.r[0].z - m.r[2].x) * s;
        t[2] = (m.r[1].x - m.r[0].y) * s;
    } else {
        int i = m.r[0].x < m.r[1].y ? (m.r[1].y < m.r[2].z ? 2 : 1) : (m.r[0].x < m.r[2].z ? 2 : 0);
        int j = (i + 1) % 3;
        int k = (i + 2) % 3;
        float s = sqrtf(m.r[i][i] - m.r[j][j] - m.r[k][k] + 1.0f);
        t[i] = s * 0.5f;
        s = 0.5f / s;
        t[3] = (m.r[k][j] - m.r[j][k]) * s;
        t[j] = (m.r[j][i] + m.r[i][j]) * s;
        t[k] = (m.r[k][i] + m.r[i][k]) * s;
    }
    return Quat(t[0], t[1], t[2], t[3]);
}
// btMatrix3x3::setRotation
RL_HDI M3 quat_to_mat(Quat q) {
    float d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    float s = 2.0f / d;
    float xs = q.x * s, ys = q.y * s, zs = q.z * s;
    float wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    float xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
    float yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    return M3(V3(1.0f - (yy + zz), xy - wz, xz + wy), V3(xy + wz, 1.0f - (xx + zz), yz - wx),
              V3(xz - wy, yz + wx, 1.0f - (xx + yy)));
}
// btQuaternion(axis, angle) -> matrix; used for the wheel steering transform
RL_HDI Quat quat_axis_angle(V3 axis, float angle) {
    float d = len(axis);
    float s = rl_sin(angle * 0.5f) / d;
    return Quat(axis.x * s, axis.y * s, axis.z * s, rl_cos(angle * 0.5f));
}

// btMatrix3x3::setEulerYPR(yaw, pitch, roll) == setEulerZYX(roll, pitch, yaw)
RL_HD RL_NOINLINE inline M3 euler_ypr_to_mat(float yaw, float pitch, float roll) {
    float eulerX = roll, eulerY = pitch, eulerZ = yaw;
    float ci = rl_cos(eulerX), cj = rl_cos(eulerY), ch = rl_cos(eulerZ);
    float si = rl_sin(eulerX), sj = rl_sin(eulerY), sh = rl_sin(eulerZ);
    float cc = ci * ch, cs = ci * sh, sc = si * ch, ss = si * sh;
    return M3(V3(cj * ch, sj * sc - cs, sj * cc + ss), V3(cj * sh, sj * ss + cc, sj * cs - sc),
              V3(-sj, cj * si, cj * ci));
}
// Angle(yaw,pitch,roll).ToRotMat() (reference MathTypes.cpp:73-78)
RL_HDI M3 angle_to_rotmat(float yaw, float pitch, float roll) { return euler_ypr_to_mat(yaw, -pitch, -roll); }

// ---- reference Vec helpers (R/Math/MathTypes/MathTypes.h) — they include the zero 4th lane ----
// (strict: these feed the bit-exact rewards; out of line so the IEEE division / square-root sequences exist once)
RL_HDI float ref_dot(V3 a, V3 b) { return s_add(s_add(s_add(s_mul(a.x, b.x), s_mul(a.y, b.y)), s_mul(a.z, b.z)), 0.f); }
RL_HDI V3 s_sub3(V3 a, V3 b) { return V3(s_sub(a.x, b.x), s_sub(a.y, b.y), s_sub(a.z, b.z)); }
RL_HD RL_NOINLINE inline float ref_len(V3 v) {
    float l2 = ref_dot(v, v);
    return l2 > 0 ? s_sqrt(l2) : 0.f;
}
RL_HD RL_NOINLINE inline V3 ref_normalized(V3 v) {
    float l = ref_len(v);
    if (l > kEps * kEps) return V3(s_div(v.x, l), s_div(v.y, l), s_div(v.z, l));
    return V3(0, 0, 0);
}
RL_HDI V3 to_uu(V3 v) { return V3(s_mul(v.x, 50.f), s_mul(v.y, 50.f), s_mul(v.z, 50.f)); }

RL_HD RL_NOINLINE inline float rl_sin(float x) { return sinf(x); }
RL_HD RL_NOINLINE inline float rl_cos(float x) { return cosf(x); }
RL_HD RL_NOINLINE inline float rl_atan2(float y, float x) { return atan2f(y, x); }
RL_HD RL_NOINLINE inline float rl_asin(float x) { return asinf(x); }

// piecewise-linear curves (reference Math.cpp:7-38 LinearPieceCurve::GetOutput)
template <int N>
RL_HDI float curve(const float (&xs)[N], const float (&ys)[N], float in) {
    if (in <= xs[0]) return ys[0];
    for (int i = 1; i < N; i++) {
        if (xs[i] > in) {
            float range = xs[i] - xs[i - 1];
            float diff = ys[i] - ys[i - 1];
            float f = (in - xs[i - 1]) / range;
            return ys[i - 1] + diff * f;
        }
    }
    return ys[N - 1];
}

}  // namespace rl
