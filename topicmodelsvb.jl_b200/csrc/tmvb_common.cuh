// tmvb_common.cuh -- shared host/device helpers for libtmvb (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/tmvb.h"

namespace tmvb {

// ---------------------------------------------------------------- errors ------------------
std::string &last_error();
int fail(int code, const char *fmt, ...);

#define TMVB_CUDA(expr)                                                                        \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return ::tmvb::fail((int)_e, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,  \
                                cudaGetErrorString(_e));                                       \
    } while (0)

#define TMVB_CHECK_ARG(cond, msg)                                                              \
    do {                                                                                       \
        if (!(cond)) return ::tmvb::fail(-1, "invalid argument: %s (%s)", msg, #cond);         \
    } while (0)

#define TMVB_TRY(expr)                                                                         \
    do {                                                                                       \
        int _r = (expr);                                                                       \
        if (_r != 0) return _r;                                                                \
    } while (0)

// ---------------------------------------------------------------- constants ---------------
// EPSILON = eps(1e-14) = 2^-99 (utils.jl:3); representable in fp32 (min normal 2^-126).
#define TMVB_EPS 1.5777218104420236e-30f
#define TMVB_EPS_D 1.5777218104420236e-30

constexpr int kWarp = 32;
constexpr int kLanesPerToken = 8;  // lanes that share one token's K-vector

// ---------------------------------------------------------------- device math -------------
#ifdef __CUDACC__

__device__ __forceinline__ float warp_xor_add(float v, int m) { return v + __shfl_xor_sync(0xffffffffu, v, m); }

// sum over the 8 lanes that share a token (lane bits 0..2)
__device__ __forceinline__ float group8_sum(float v)
{
    v = warp_xor_add(v, 1);
    v = warp_xor_add(v, 2);
    v = warp_xor_add(v, 4);
    return v;
}
// sum over the 4 token streams of a warp (lane bits 3..4)
__device__ __forceinline__ float streams_sum(float v)
{
    v = warp_xor_add(v, 8);
    v = warp_xor_add(v, 16);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) { return streams_sum(group8_sum(v)); }
__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// psi(x) and ln Gamma(x) for x > 0 in fp32, evaluated in-register.
//
// For x < 6 the argument is shifted by 6 with the recurrence folded into one rational term:
//   P(x) = x(x+1)...(x+5),  psi(x) = psi(x+6) - P'(x)/P(x),  lnG(x) = lnG(x+6) - ln P(x)
// (one division instead of the reference's six, utils.jl:27-36), then the asymptotic series
//   psi(y) ~ ln y - 1/(2y) - 1/(12y^2) + 1/(120y^4) - 1/(252y^6)          (|err| < 3e-9, y >= 6)
//   lnG(y) ~ (y-.5) ln y - y + .5 ln(2pi) + 1/(12y) - 1/(360y^3) + 1/(1260y^5)
// replace the reference's 8-term fp32 series (utils.jl:39-50; Koelbig 1972).
struct PsiLg {
    float psi, lg;
};

template <bool WANT_LG, bool FAST = false>
__device__ __forceinline__ PsiLg psi_lgamma(float x)
{
    float P = 1.0f, D = 0.0f, y = x;
    if (x < 6.0f) {
        // P = prod (x+k), D = dP/dx via the product rule, k = 0..5
        P = x;
        D = 1.0f;
#pragma unroll
        for (int k = 1; k < 6; k++) {
            float f = x + (float)k;
            D = fmaf(D, f, P);
            P *= f;
        }
        y = x + 6.0f;
    }
    float ly = FAST ? __logf(y) : logf(y);
    float t = FAST ? __fdividef(1.0f, y) : __frcp_rn(y), t2 = t * t;  // FAST: MUFU.RCP (1 ulp) instead of the IEEE sequence
    PsiLg r;
    float ser = t2 * (8.3333333333e-2f - t2 * (8.3333333333e-3f - t2 * 3.9682539683e-3f));
    r.psi = ly - 0.5f * t - ser;
    if (x < 6.0f) r.psi -= FAST ? __fdividef(D, P) : D / P;
    r.lg = 0.0f;
    if (WANT_LG) {
        float sl = t * (8.3333333333e-2f - t2 * (2.7777777778e-3f - t2 * 7.9365079365e-4f));
        r.lg = fmaf(y - 0.5f, ly, -y) + 0.91893853320467f + sl;
        if (x < 6.0f) r.lg -= logf(P);
    }
    return r;
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2: two lanes of fp32 per 64-bit register pair) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float hsum2(f32x2 v)
{
    float lo, hi;
    unpk2(v, lo, hi);
    return lo + hi;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
// 4-byte asynchronous global -> shared copy (LDGSTS.32)
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier transaction tracking ------------
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(s), "r"(parity)
        : "memory");
}
// order earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// one contiguous global -> shared copy of `bytes` (multiple of 16; both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d), "l"(gmem_src),
                 "r"(bytes), "r"(b)
                 : "memory");
}

// fast fp32 transcendental forms used inside the per-sweep K phase
__device__ __forceinline__ float fast_log(float x) { return __logf(x); }   // MUFU.LG2 * ln2
__device__ __forceinline__ float fast_exp(float x) { return __expf(x); }   // MUFU.EX2(x * log2e)

// c / s for s > 0 well inside the normal range (s >= K * 2^-99 by construction): one MUFU.RCP + one FMUL, without the
// range-reduction guard __fdividef carries (6 more instructions per token round)
__device__ __forceinline__ float fast_div_pos(float c, float s)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    return c * r;
}

// fire-and-forget fp32 add into global memory (RED.E.ADD.F32)
__device__ __forceinline__ void red_add(float *addr, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(addr), "f"(v) : "memory");
}

// 16-byte vector form (REDG.E.ADD.F32x4): addr must be 16-byte aligned
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

#endif  // __CUDACC__

// ---------------------------------------------------------------- host fp64 special functions
// rank the vocabulary per topic on the device (tmvb_sort.cu); out[i*V + r] = 1-based term id of rank r
int topics_argsort(const float *d_mat, const float *d_scale, int K, int ld, int V, int *d_out, void **ws, size_t *ws_bytes,
                   cudaStream_t stream, int n_sm);

double h_digamma(double x);
double h_trigamma(double x);

}  // namespace tmvb
