// tmvb_filt.cuh -- what the two filtered models (fLDA.jl, fCTM.jl) share on the device: the token pass of
//     phi_ni = softmax_i(tau_n ln(beta[i, w_n] + eps) + E_i)          update_phi!  (fLDA.jl:198-201, fCTM.jl:239-242)
// in base 2 over a K x V table L = log2(beta + eps) (see tmvb_flda.cu), and the host-side launches of the kappa / table / tau
// plumbing kernels that live in tmvb_flda.cu.
#pragma once

#include "tmvb_shard.cuh"

namespace tmvb {

// L[which] = log2(beta + eps) (pad topics 0) for the n = V * K_ld entries of a table
int filt_log_table(Shard *s, const float *beta, float *L);
// update_kappa!(model) (fLDA.jl:149-153 / fCTM.jl:154-158): kappa_old <- kappa; kappa = kstats ./ sum(kstats); kstats <- 0
int filt_kappa_update(Shard *s, float *kstats, float *kappa, float *kappa_old);
// kq[j] = (1 - eta) kappa[j] (j < V), kq[V] = eta
int filt_push_kq(Shard *s, const float *kappa, float *kq, double eta);
// per-token arrays between the caller's CSR order and the shard's internal order; the upload checks 0 <= tau <= 1 (*bad = 1 otherwise)
int filt_tau_upload(Shard *s, const float *host_tau, float *d_tau, int *bad);
int filt_tau_download(Shard *s, const float *d_tau, float *host_tau);

#ifdef __CUDACC__

constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
constexpr float kPadE = -1000.0f;   // E2 of a pad topic: 2^-1000 flushes to zero

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

// p = 2^(tau L + E2) for this lane's chunks of one row; returns the lane's partial s and q (packed pairs summed by the caller)
template <int CPL>
__device__ __forceinline__ void flda_row(const ulonglong2 (&b)[CPL], const f32x2 (&E01)[CPL], const f32x2 (&E23)[CPL], float tau, f32x2 (&p01)[CPL],
                                         f32x2 (&p23)[CPL], float &s, float &q)
{
    const f32x2 t2 = pk2(tau, tau);
    f32x2 sa = 0ull, sb = 0ull, qa = 0ull, qb = 0ull;
#pragma unroll
    for (int m = 0; m < CPL; m++) {
        float x0, x1, x2, x3;
        unpk2(fma2(t2, b[m].x, E01[m]), x0, x1);
        unpk2(fma2(t2, b[m].y, E23[m]), x2, x3);
        p01[m] = pk2(ex2_ftz(x0), ex2_ftz(x1));
        p23[m] = pk2(ex2_ftz(x2), ex2_ftz(x3));
        sa = add2(sa, p01[m]);
        sb = add2(sb, p23[m]);
        qa = fma2(p01[m], b[m].x, qa);
        qb = fma2(p23[m], b[m].y, qb);
    }
    s = hsum2(add2(sa, sb));
    q = hsum2(add2(qa, qb));
}

template <int LPT, int CPL>
__device__ __forceinline__ void flda_load_row(const float *tile, const float *gL, int RS, int K_ld, int CH, int n, int cap, int term, int kl,
                                              ulonglong2 (&b)[CPL])
{
    const ulonglong2 zero = make_ulonglong2(0ull, 0ull);
    if (n < cap) {
        const ulonglong2 *row = reinterpret_cast<const ulonglong2 *>(tile + (size_t)n * RS) + kl;
#pragma unroll
        for (int m = 0; m < CPL; m++) b[m] = (kl + LPT * m < CH) ? row[LPT * m] : zero;
    } else {
        const ulonglong2 *row = reinterpret_cast<const ulonglong2 *>(gL + (size_t)term * K_ld) + kl;
#pragma unroll
        for (int m = 0; m < CPL; m++) b[m] = (kl + LPT * m < CH) ? __ldg(row + LPT * m) : zero;
    }
}

#endif  // __CUDACC__

}  // namespace tmvb
