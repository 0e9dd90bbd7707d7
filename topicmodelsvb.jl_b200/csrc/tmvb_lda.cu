// tmvb_lda.cu -- LDA coordinate-ascent VB on sm_100a: fused per-document E-step + statistic
// scatter, M-step normalisation, ELBO, and the tmvb_lda_* C ABI (include/tmvb.h).
//
// Reference semantics followed: the CPU model src/LDA.jl (per-document stopping rule, lagged-phi
// ELBO); the thing replaced: src/gpuLDA.jl's 7 OpenCL kernels + modelutils.jl:370-397,501-516.
//
// Device data layout (all owned by the handle):
//   beta[2]      float [V][K_ld]   term-major rows ("term rows", == Julia's column-major K x V with
//                                  the leading dimension padded to K_ld = 8*ceil(K/8), pad = 0); double
//                                  buffered so beta_old (LDA.jl:122) costs nothing
//   stats        float [V][K_ld]   sufficient statistics beta_temp (LDA.jl:131), RED.ADD target
//   Elogtheta, Elogtheta_old, gamma  float [M][K_ld]   per-document K-vectors, internal doc order
//   doc_off int64 [M+1], terms int32 [nnz], counts float [nnz]   CSR, documents sorted by length
//                                  (descending) so that equal-sized documents share a launch
//   small        double [K_ld+2]   sum_d Elogtheta_d | per-document ELBO terms | sweep counter
//                                  (summed across ranks together with stats in multi-GPU runs)
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "tmvb_common.cuh"

namespace tmvb {

struct LdaDev {
    int K, K_ld, V, RS;
    long long M;
    const float *beta;
    const float *alpha;
    float *stats;
    const long long *doc_off;
    const int *terms;
    const float *counts;
    float *Elogtheta, *Elogtheta_old, *gamma;
    double *small;
    int viter;
    float vtol;
    int stage_bulk;  // 1: TMA bulk row copies (UBLKCP), 0: 16-byte cp.async (LDGSTS)
    int dbg;         // developer probes: bit0 skip the scatter, bit1 skip the final pass
};

// Thread mapping of the E-step: ONE WARP PER DOCUMENT (a CTA is a single warp; the grid is
// persistent and warps pull documents, longest first, from an atomic counter).
//   token phase  -- lane (ts = lane / LPT, kl = lane % LPT) owns token stream ts (S = 32/LPT streams)
//                   and the 16-byte chunks q = kl + LPT*m (m < CPL) of every term row it visits,
//                   i.e. topics 4q..4q+3, read from the shared-memory tile with LDS.128.
//   K phase      -- lane l owns topics i = l + 32r (r < R): gamma, digamma, exp and the
//                   convergence test are evaluated once per sweep with 1-2 topics per lane.
// The two layouts meet in shared memory: per-stream partial K-vectors are written as float4
// chunks (gs), summed by the owner lanes, and exp(Elogtheta) travels back through e_s.
// (LPT, CPL) is chosen per K on the host (lda_pick_layout); the row stride RS of tile/gs is padded
// so that RS/4 = LPT (mod 2 LPT) for LPT < 8, which makes every LDS.128/STS.128 phase conflict-free.
__host__ __device__ inline size_t lda_smem_bytes(int RS, int LPT, int cap)
{
    size_t b = (size_t)cap * RS * 4 + (size_t)cap * 8;  // tile + counts + terms
    b += (size_t)(32 / LPT) * RS * 4;                   // gs: per-stream partial K-vectors
    b += (size_t)RS * 4;                                // e_s: exp(Elogtheta)
    b += 16;                                            // mbarrier of the TMA staging
    return b;
}

template <int LPT>
__device__ __forceinline__ float group_sum(float v)
{
#pragma unroll
    for (int m = 1; m < LPT; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
template <int LPT>
__device__ __forceinline__ float across_streams_sum(float v)
{
#pragma unroll
    for (int m = LPT; m < 32; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// One pass over the document's tokens.
// Pass 1 (per token n):  s_n = K*eps + sum_i beta[i,w_n] e_i ;  t_n = c_n / s_n
// Pass 2:                g_i += beta[i,w_n] t_n   so that   (phi * counts)_i = e_i g_i + eps sum_n t_n
// which is update_phi! + update_gamma! (LDA.jl:143-154) without ever forming phi.
// FINAL instead scatters c_n phi_ni = t_n (eps + beta e_i) into stats (LDA.jl:129-132) with
// 16-byte vector reductions and accumulates sum_n c_n H(phi_n) (LDA.jl:76-80).
template <int LPT, int CPL, bool OVF, bool FINAL, bool ELBO>
__device__ __forceinline__ void lda_token_pass(const LdaDev &p, const float *tile, const float *cnt_s,
                                               const int *term_s, long long o, int Nd, int cap, int rounds,
                                               int ts, int kl, const float4 (&e)[CPL], float4 (&g)[CPL],
                                               float &tsum, float &ent)
{
    constexpr int S = 32 / LPT;
    const int K_ld = p.K_ld, CH = K_ld >> 2, RS = p.RS;
    const float Keps = (float)p.K * TMVB_EPS;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int r = 0; r < rounds; r++) {
        const int n = r * S + ts;
        const bool ok = n < Nd;
        float4 b[CPL];
        float c = 0.0f;
        int term = 0;
        if (!OVF || n < cap) {
            const int nn = ok ? n : 0;
            const float4 *row = reinterpret_cast<const float4 *>(tile + nn * RS) + kl;
#pragma unroll
            for (int m = 0; m < CPL; m++) b[m] = (m < CPL - 1 || kl + LPT * m < CH) ? row[LPT * m] : zero4;
            if (ok) c = cnt_s[nn];
            if (FINAL) term = term_s[nn];
        } else {
            const long long q = o + (ok ? n : 0);
            term = p.terms[q];
            const float4 *row = reinterpret_cast<const float4 *>(p.beta + (size_t)term * K_ld) + kl;
#pragma unroll
            for (int m = 0; m < CPL; m++) b[m] = (m < CPL - 1 || kl + LPT * m < CH) ? __ldg(row + LPT * m) : zero4;
            if (ok) c = p.counts[q];
        }
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
        for (int m = 0; m < CPL; m++) {
            s0 = fmaf(b[m].x, e[m].x, s0);
            s1 = fmaf(b[m].y, e[m].y, s1);
            s2 = fmaf(b[m].z, e[m].z, s2);
            s3 = fmaf(b[m].w, e[m].w, s3);
        }
        const float s = group_sum<LPT>((s0 + s1) + (s2 + s3)) + Keps;
        const float t = __fdividef(c, s);
        if (!FINAL) {
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                g[m].x = fmaf(b[m].x, t, g[m].x);
                g[m].y = fmaf(b[m].y, t, g[m].y);
                g[m].z = fmaf(b[m].z, t, g[m].z);
                g[m].w = fmaf(b[m].w, t, g[m].w);
            }
            tsum += t;
        } else if (ok) {
            float *srow = p.stats + (size_t)term * K_ld + 4 * kl;
            float a = 0.0f;
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                const int i0 = 4 * (kl + LPT * m);
                if (i0 < p.K) {
                    // pad topics (i >= K) carry beta = e = 0: they get t*eps, which the M-step ignores
                    const float ux = fmaf(b[m].x, e[m].x, TMVB_EPS), uy = fmaf(b[m].y, e[m].y, TMVB_EPS);
                    const float uz = fmaf(b[m].z, e[m].z, TMVB_EPS), uw = fmaf(b[m].w, e[m].w, TMVB_EPS);
                    if (!(p.dbg & 1)) red_add_v4(srow + 4 * LPT * m, t * ux, t * uy, t * uz, t * uw);
                    if (ELBO) {
                        a = fmaf(t * ux, __logf(ux), a);
                        if (i0 + 1 < p.K) a = fmaf(t * uy, __logf(uy), a);
                        if (i0 + 2 < p.K) a = fmaf(t * uz, __logf(uz), a);
                        if (i0 + 3 < p.K) a = fmaf(t * uw, __logf(uw), a);
                    }
                }
            }
            if (ELBO) ent += ((kl == 0) ? c * __logf(s) : 0.0f) - a;
        }
    }
}

template <int LPT, int CPL, bool ELBO>
__global__ void __launch_bounds__(32) lda_estep_kernel(const LdaDev p, int doc_begin, int doc_end, int cap, int *counter)
{
    constexpr int S = 32 / LPT;                   // token streams per warp
    constexpr int R = (LPT * CPL + 7) / 8;        // K-phase topics per lane (>= ceil(K_ld / 32))
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int kl = lane % LPT, ts = lane / LPT;
    const int K = p.K, K_ld = p.K_ld, CH = K_ld >> 2, RS = p.RS;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw);
    float *tile = reinterpret_cast<float *>(smem_raw + 16);
    float *gs = tile + (size_t)cap * RS;                       // [S][RS]
    float *e_s = gs + (size_t)S * RS;                          // [RS]
    float *cnt_s = e_s + RS;                                   // [cap]
    int *term_s = reinterpret_cast<int *>(cnt_s + cap);        // [cap]

    // K-phase state: topics i = lane + 32 r
    float alpha_k[R], Eold_k[R], Enew_k[R], e_k[R], gam_k[R];
    double esum_k[R];
    float asum = 0.0f;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = lane + 32 * r;
        alpha_k[r] = (i < K) ? p.alpha[i] : 0.0f;
        asum += alpha_k[r];
        esum_k[r] = 0.0;
        Enew_k[r] = gam_k[r] = Eold_k[r] = e_k[r] = 0.0f;
    }
    asum = warp_sum(asum);
    // convergence test in fixed point: sum_i (dE_i)^2 * (2^20 / vtol^2) < 2^20, summed with one REDUX
    const float dscale = (p.vtol > 0.0f) ? 1048576.0f / (p.vtol * p.vtol) : 0.0f;
    double elbo_thr = 0.0;
    unsigned long long sweeps_thr = 0;
    unsigned phase = 0;
    if (p.stage_bulk) {
        if (lane == 0) mbar_init(mbar, 1);
        __syncwarp();
    }

    for (;;) {
        int d = 0;
        if (lane == 0) d = doc_begin + atomicAdd(counter, 1);
        d = __shfl_sync(0xffffffffu, d, 0);
        if (d >= doc_end) break;
        const long long o = p.doc_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o);
        const int ns = min(Nd, cap);
        const bool ovf = Nd > cap;
        const int rounds = (Nd + S - 1) / S;

        // stage the document: term ids + counts, then its K x N_d slab of beta -- one TMA bulk copy
        // per term row (a row is K_ld*4 contiguous bytes in HBM/L2), completion tracked by an mbarrier
        __syncwarp();
        float csum = 0.0f;
        for (int n = lane; n < Nd; n += 32) {
            const float c = p.counts[o + n];
            csum += c;
            if (n < ns) {
                term_s[n] = p.terms[o + n];
                cnt_s[n] = c;
            }
        }
        if (p.stage_bulk) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(mbar, (unsigned)(ns * K_ld * 4));
            __syncwarp();
            for (int n = lane; n < ns; n += 32)
                bulk_g2s(tile + n * RS, p.beta + (size_t)term_s[n] * K_ld, (unsigned)(K_ld * 4), mbar);
        } else {
            __syncwarp();
            for (int c = lane; c < ns * CH; c += 32) {
                const int n = c / CH, q = c - n * CH;
                cp_async16(tile + n * RS + 4 * q, p.beta + (size_t)term_s[n] * K_ld + 4 * q);
            }
            cp_async_commit();
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            Eold_k[r] = (i < K) ? p.Elogtheta[(size_t)d * K_ld + i] : 0.0f;
            e_k[r] = (i < K) ? expf(Eold_k[r]) : 0.0f;
            if (i < K_ld) e_s[i] = e_k[r];
        }
        // sum(gamma_d) = sum(alpha) + sum_n c_n + K*EPS whatever phi is (each phi column sums to one),
        // so digamma(sum gamma) (LDA.jl:138) is a per-document constant
        const float gsum = (asum + warp_sum(csum)) + (float)K * TMVB_EPS;
        const float psi_sum = psi_lgamma<false>(gsum).psi;
        if (p.stage_bulk) {
            mbar_wait(mbar, phase);
            phase ^= 1u;
        } else {
            cp_async_wait_all();
        }
        __syncwarp();

        float4 e[CPL];
        int v = 0;
        for (;;) {
            // ---- token phase
#pragma unroll
            for (int m = 0; m < CPL; m++)
                e[m] = (m < CPL - 1 || kl + LPT * m < CH) ? reinterpret_cast<const float4 *>(e_s)[kl + LPT * m] : zero4;
            float4 g[CPL];
            float tsum = 0.0f, dummy = 0.0f;
#pragma unroll
            for (int m = 0; m < CPL; m++) g[m] = zero4;
            if (!ovf)
                lda_token_pass<LPT, CPL, false, false, false>(p, tile, cnt_s, term_s, o, Nd, cap, rounds, ts, kl, e, g, tsum, dummy);
            else
                lda_token_pass<LPT, CPL, true, false, false>(p, tile, cnt_s, term_s, o, Nd, cap, rounds, ts, kl, e, g, tsum, dummy);
#pragma unroll
            for (int m = 0; m < CPL; m++)
                if (m < CPL - 1 || kl + LPT * m < CH) reinterpret_cast<float4 *>(gs + ts * RS)[kl + LPT * m] = g[m];
            const float tt = across_streams_sum<LPT>(tsum);
            __syncwarp();

            // ---- K phase: update_gamma! (LDA.jl:143-146), update_Elogtheta! (LDA.jl:136-139)
            float dpart = 0.0f;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = lane + 32 * r;
                float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f, g3 = 0.0f;
                if (i < K) {
                    if (S >= 4) {
#pragma unroll
                        for (int w = 0; w < S; w += 4) {
                            g0 += gs[w * RS + i];
                            g1 += gs[(w + 1) * RS + i];
                            g2 += gs[(w + 2) * RS + i];
                            g3 += gs[(w + 3) * RS + i];
                        }
                    } else {
#pragma unroll
                        for (int w = 0; w < S; w++) g0 += gs[w * RS + i];
                    }
                }
                const float gi = (g0 + g1) + (g2 + g3);
                // gamma = EPS + (alpha + phi*counts),   phi*counts = e .* g + eps * sum_n t_n
                gam_k[r] = (i < K) ? (alpha_k[r] + fmaf(e_k[r], gi, TMVB_EPS * tt)) + TMVB_EPS : 1.0f;
                Enew_k[r] = psi_lgamma<false, true>(gam_k[r]).psi - psi_sum;
                if (i < K) {
                    const float df = Enew_k[r] - Eold_k[r];
                    dpart = fmaf(df, df, dpart);
                }
            }
            v++;
            // LDA.jl:175: stop when ||Elogtheta - Elogtheta_old||_2 < vtol (or after viter sweeps)
            if (v >= p.viter) break;
            // per-lane clamp 2^26 keeps the 32-lane integer sum below 2^31
            if (dscale > 0.0f && __reduce_add_sync(0xffffffffu, (unsigned)fminf(dpart * dscale, 67108864.0f)) < 1048576u) break;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = lane + 32 * r;
                Eold_k[r] = Enew_k[r];
                e_k[r] = (i < K) ? fast_exp(Enew_k[r]) : 0.0f;
                if (i < K_ld) e_s[i] = e_k[r];
            }
            __syncwarp();
        }

        // update_beta!(model, d) (LDA.jl:129-132): scatter the last phi, weighted by counts
        if (!(p.dbg & 2)) {
            float4 g[CPL];
            float tsum = 0.0f, ent = 0.0f;
            if (!ovf)
                lda_token_pass<LPT, CPL, false, true, ELBO>(p, tile, cnt_s, term_s, o, Nd, cap, rounds, ts, kl, e, g, tsum, ent);
            else
                lda_token_pass<LPT, CPL, true, true, ELBO>(p, tile, cnt_s, term_s, o, Nd, cap, rounds, ts, kl, e, g, tsum, ent);
            if (ELBO) elbo_thr += (double)ent;
        }

        float a = 0.0f;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            if (i < K_ld) {
                const bool ok = i < K;
                p.gamma[(size_t)d * K_ld + i] = ok ? gam_k[r] : 0.0f;
                p.Elogtheta[(size_t)d * K_ld + i] = ok ? Enew_k[r] : 0.0f;
                p.Elogtheta_old[(size_t)d * K_ld + i] = ok ? Eold_k[r] : 0.0f;
                if (ok) {
                    esum_k[r] += (double)Enew_k[r];
                    if (ELBO) a += psi_lgamma<true>(gam_k[r]).lg;
                }
            }
        }
        // Dirichlet entropy (utils.jl:163-180) + Elogpz (LDA.jl:57-60): with gamma = alpha + phi*c
        // and psi(gamma_i) = Elogtheta_i + psi(sum gamma) they collapse to
        //   sum_i lnG(gamma_i) - lnG(sum gamma) + sum_i (1 - alpha_i) Elogtheta_i ;
        // the last sum is linear in sum_d Elogtheta_d and is added on the host in fp64.
        if (ELBO) {
            if (lane == 0) a -= psi_lgamma<true>(gsum).lg;
            elbo_thr += (double)a;
        }
        if (lane == 0) sweeps_thr += (unsigned long long)v;
    }

    // flush the warp's accumulators
    if (ELBO) {
        const double tot = warp_sum_d(elbo_thr);
        if (lane == 0 && tot != 0.0) atomicAdd(p.small + K_ld, tot);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = lane + 32 * r;
        if (i < K && esum_k[r] != 0.0) atomicAdd(p.small + i, esum_k[r]);
    }
    if (lane == 0 && sweeps_thr) atomicAdd(p.small + K_ld + 1, (double)sweeps_thr);
}

// ------------------------------------------------------------------ M-step ------------------
// rowsum_i = sum_j stats[j][i]   (the `sum(beta_temp, dims=2)` of LDA.jl:123)
__global__ void lda_colsum_kernel(const float *__restrict__ stats, int V, int K_ld, double *__restrict__ rowsum)
{
    extern __shared__ double sh[];
    const int R = blockDim.x / K_ld;
    const int r = threadIdx.x / K_ld, i = threadIdx.x - r * K_ld;
    double acc = 0.0;
    if (r < R)
        for (int j = blockIdx.x * R + r; j < V; j += gridDim.x * R) acc += (double)stats[(size_t)j * K_ld + i];
    sh[threadIdx.x] = (r < R) ? acc : 0.0;
    __syncthreads();
    if (threadIdx.x < K_ld) {
        double a = 0.0;
        for (int q = 0; q < R; q++) a += sh[q * K_ld + threadIdx.x];
        if (a != 0.0) atomicAdd(rowsum + threadIdx.x, a);
    }
}

// beta_new = stats ./ rowsum ; stats <- 0 ; elbo_w += sum stats * ln(beta_new + eps)
// (LDA.jl:121-125 and the Elogpw term LDA.jl:64-67 rewritten over the statistics)
__global__ void lda_normalize_kernel(float *__restrict__ stats, float *__restrict__ beta_new, const double *__restrict__ rowsum,
                                     long long n, int K, int K_ld, double *__restrict__ elbo_w, int want_elbo)
{
    double acc = 0.0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q % K_ld);
        float s = stats[q];
        float b = 0.0f;
        if (i < K) {
            const double rs = rowsum[i];
            b = rs > 0.0 ? (float)((double)s / rs) : 0.0f;
            if (want_elbo) acc += (double)(s * logf(b + TMVB_EPS));
        }
        beta_new[q] = b;
        stats[q] = 0.0f;
    }
    if (want_elbo) {
        acc = warp_sum_d(acc);
        if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(elbo_w, acc);
    }
}

// ------------------------------------------------------------------ standalone ELBO ---------
__device__ inline double d_digamma(double x)
{
    double r = 0.0;
    while (x < 10.0) {
        r -= 1.0 / x;
        x += 1.0;
    }
    double t = 1.0 / x, t2 = t * t;
    double s = t2 * (1.0 / 12 - t2 * (1.0 / 120 - t2 * (1.0 / 252 - t2 * (1.0 / 240 - t2 * (1.0 / 132 - t2 * (691.0 / 32760 - t2 * (1.0 / 12)))))));
    return r + log(x) - 0.5 * t - s;
}

// update_elbo! exactly as the CPU model states it (LDA.jl:50-93): phi rebuilt from beta_old and
// Elogtheta_old, the five expectations evaluated with alpha, beta, gamma, Elogtheta.  fp64
// arithmetic on the fp32 device state; one warp per document, lanes over topics.
template <typename real>
__global__ void lda_elbo_kernel(const LdaDev p, const float *__restrict__ beta_old, double lg_alpha_term, double *out)
{
    constexpr bool F64 = sizeof(real) == 8;
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const real eps = (real)TMVB_EPS_D;
    double acc = 0.0;
    for (long long d = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); d < p.M; d += (long long)gridDim.x * wpb) {
        const long long o = p.doc_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o);
        const float *En = p.Elogtheta + d * p.K_ld, *Eo = p.Elogtheta_old + d * p.K_ld, *gm = p.gamma + d * p.K_ld;
        double dacc = 0.0, g0 = 0.0;
        for (int i = lane; i < p.K; i += 32) {
            const double g = gm[i], E = En[i];
            g0 += g;
            if (F64) {
                dacc += ((double)p.alpha[i] - 1.0) * E + lgamma(g) - (g - 1.0) * d_digamma(g);
            } else {
                const PsiLg pl = psi_lgamma<true>((float)g);
                dacc += ((double)p.alpha[i] - 1.0) * E + (double)pl.lg - (g - 1.0) * (double)pl.psi;
            }
        }
        g0 = warp_sum_d(g0);
        real tacc = 0;
        for (int n = 0; n < Nd; n++) {
            const int term = p.terms[o + n];
            const real c = p.counts[o + n];
            const float *bo = beta_old + (size_t)term * p.K_ld, *bn = p.beta + (size_t)term * p.K_ld;
            real s = 0;
            for (int i = lane; i < p.K; i += 32) s += eps + (real)bo[i] * (F64 ? (real)exp((double)Eo[i]) : (real)expf(Eo[i]));
            s = F64 ? (real)warp_sum_d((double)s) : (real)warp_sum((float)s);
            real a = 0;
            for (int i = lane; i < p.K; i += 32) {
                const real u = eps + (real)bo[i] * (F64 ? (real)exp((double)Eo[i]) : (real)expf(Eo[i]));
                const real ph = u / s;
                const real lb = F64 ? (real)log((double)bn[i] + TMVB_EPS_D) : (real)logf(bn[i] + TMVB_EPS);
                const real lp = F64 ? (real)(ph > 0 ? log((double)ph) : 0.0) : (real)(ph > 0 ? logf((float)ph) : 0.f);
                a += ph * ((real)En[i] + lb - lp);
            }
            if (F64) dacc += (double)(c * a); else tacc += c * a;
        }
        dacc += (double)tacc;
        dacc = warp_sum_d(dacc);
        if (lane == 0) {
            double ent = 0.0;
            if (p.K > 1) ent = -lgamma(g0) + (g0 - (double)p.K) * d_digamma(g0);
            acc += dacc + ent + lg_alpha_term;
        }
    }
    if (lane == 0 && acc != 0.0) atomicAdd(out, acc);
}

// phi[K x sumN] in the caller's token order, rebuilt from beta_old / Elogtheta_old (LDA.jl:87-88)
__global__ void lda_phi_kernel(const LdaDev p, const float *__restrict__ beta_old, const long long *__restrict__ src_off, float *__restrict__ phi)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (long long d = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); d < p.M; d += (long long)gridDim.x * wpb) {
        const long long o = p.doc_off[d], so = src_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o);
        const float *Eo = p.Elogtheta_old + d * p.K_ld;
        for (int n = 0; n < Nd; n++) {
            const float *bo = beta_old + (size_t)p.terms[o + n] * p.K_ld;
            float s = 0.0f;
            for (int i = lane; i < p.K; i += 32) s += fmaf(bo[i], expf(Eo[i]), TMVB_EPS);
            s = warp_sum(s);
            for (int i = lane; i < p.K; i += 32) phi[(size_t)(so + n) * p.K + i] = fmaf(bo[i], expf(Eo[i]), TMVB_EPS) / s;
        }
    }
}

// ------------------------------------------------------------------ layout kernels ----------
// dst[p][0..K_ld) = src[perm ? perm[p] : p][0..K) , zero padded
__global__ void pad_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, const int *__restrict__ perm,
                                long long rows, int K, int K_ld)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < rows * K_ld; q += (long long)gridDim.x * blockDim.x) {
        const long long r = q / K_ld;
        const int i = (int)(q - r * K_ld);
        const long long sr = perm ? perm[r] : r;
        dst[q] = (i < K) ? src[sr * K + i] : 0.0f;
    }
}
// dst[perm ? perm[p] : p][0..K) = src[p][0..K)
__global__ void unpad_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, const int *__restrict__ perm,
                                  long long rows, int K, int K_ld)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < rows * K; q += (long long)gridDim.x * blockDim.x) {
        const long long r = q / K;
        const int i = (int)(q - r * K);
        const long long dr = perm ? perm[r] : r;
        dst[dr * K + i] = src[r * K_ld + i];
    }
}
// check_model(::gpuLDA) invariants that need a pass over the data (modelutils.jl:264-273), evaluated on
// the device copy: bit0 non-finite, bit1 sign violation (what=0: beta >= 0; 1: Elogtheta <= 0; 2: gamma > 0)
__global__ void validate_kernel(const float *__restrict__ x, long long n, int what, int *__restrict__ err)
{
    int e = 0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const float v = x[q];
        if (!isfinite(v)) e |= 1;
        if ((what == 0 && v < 0.f) || (what == 1 && v > 0.f) || (what == 2 && !(v > 0.f))) e |= 2;
    }
    if (e) atomicOr(err, e << (2 * what));
}

// CSR re-layout: internal document p takes the tokens of caller document perm[p]; Int64 -> int32 / float
__global__ void pack_corpus_kernel(const long long *__restrict__ terms64, const long long *__restrict__ counts64,
                                   const long long *__restrict__ src_off, const long long *__restrict__ dst_off, long long M,
                                   int V, int *__restrict__ terms, float *__restrict__ counts, int *__restrict__ err)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (long long d = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); d < M; d += (long long)gridDim.x * wpb) {
        const long long so = src_off[d], o = dst_off[d];
        const int Nd = (int)(dst_off[d + 1] - o);
        for (int n = lane; n < Nd; n += 32) {
            const long long t = terms64[so + n], c = counts64[so + n];
            if (t < 0 || t >= V) atomicOr(err, 1);
            if (c <= 0) atomicOr(err, 2);
            terms[o + n] = (int)t;
            counts[o + n] = (float)c;
        }
    }
}

// ------------------------------------------------------------------ host side ---------------
struct Bucket {
    int doc_begin, doc_end, cap, warps, grid;
    size_t smem;
};

typedef void (*EstepFn)(const LdaDev, int, int, int, int *);

struct Layout {
    int lpt, cpl;
    EstepFn fn[2];  // [want_elbo]
};
#define TMVB_LAYOUT(L, C) {L, C, {(EstepFn)lda_estep_kernel<L, C, false>, (EstepFn)lda_estep_kernel<L, C, true>}}
// the (LPT, CPL) pairs lda_pick_layout can select for K <= 256
static const Layout kLayouts[] = {
    TMVB_LAYOUT(1, 4), TMVB_LAYOUT(1, 8), TMVB_LAYOUT(2, 1), TMVB_LAYOUT(2, 3), TMVB_LAYOUT(2, 5), TMVB_LAYOUT(2, 6),
    TMVB_LAYOUT(2, 7), TMVB_LAYOUT(2, 8), TMVB_LAYOUT(4, 5), TMVB_LAYOUT(4, 6), TMVB_LAYOUT(4, 7), TMVB_LAYOUT(4, 8),
    TMVB_LAYOUT(8, 5), TMVB_LAYOUT(8, 6), TMVB_LAYOUT(8, 7), TMVB_LAYOUT(8, 8),
};

// row stride (in floats) of the shared-memory tile for CH 16-byte chunks per row
static int lda_row_stride(int CH, int lpt)
{
    int r = CH;
    if (lpt < 8)
        while (r % (2 * lpt) != lpt) r++;
    return 4 * r;
}

// Pick the lane layout for K: minimise (estimated warp-instructions per token) x sqrt(shared-memory inflation)
static const Layout *lda_pick_layout(int K_ld, int *RS_out)
{
    const int CH = K_ld / 4;
    const Layout *best = nullptr;
    double best_cost = 0.0;
    const int force_lpt = getenv("TMVB_LDA_LPT") ? atoi(getenv("TMVB_LDA_LPT")) : 0;
    for (const Layout &l : kLayouts) {
        if (l.lpt * l.cpl < CH) continue;
        if (force_lpt && l.lpt != force_lpt) continue;
        const int S = 32 / l.lpt, RS = lda_row_stride(CH, l.lpt);
        const double instr_tok = (l.cpl * 9.0 + 2.0 * log2((double)l.lpt) + 12.0) / S;
        const double fixed = (l.cpl + ((K_ld + 31) / 32) * 2.0 * S) / 80.0;
        const double cost = (instr_tok + fixed) * sqrt((double)RS / K_ld);
        if (!best || cost < best_cost - 1e-9) {
            best = &l;
            best_cost = cost;
            *RS_out = RS;
        }
    }
    return best;
}

}  // namespace tmvb

using namespace tmvb;

struct tmvb_lda_s {
    int device = 0, n_sm = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t K = 0, M = 0, V = 0, nnz = 0;
    size_t nnz_cap = 0;
    int K_ld = 0, RS = 0;
    const tmvb::Layout *layout = nullptr;
    bool corpus_set = false, params_set = false;
    // corpus
    long long *d_doc_off = nullptr, *d_src_off = nullptr;
    int *d_terms = nullptr, *d_perm = nullptr;
    float *d_counts = nullptr;
    std::vector<int> h_perm;
    std::vector<Bucket> buckets;
    // parameters
    float *d_alpha = nullptr, *d_beta[2] = {nullptr, nullptr}, *d_stats = nullptr;
    int cur = 0;
    float *d_Elogtheta = nullptr, *d_Elogtheta_old = nullptr, *d_gamma = nullptr;
    std::vector<double> h_alpha;  // fp64 master copy of alpha (update_alpha! runs in fp64 on the host)
    std::vector<double> h_alpha_estep;  // alpha the last E-step ran with
    // accumulators
    double *d_small = nullptr;    // [K_ld+2], summed over ranks
    double *d_local = nullptr;    // [K_ld] rowsum | [K_ld] elbo_w | [K_ld+1] scratch for mode-1 ELBO
    int *d_counters = nullptr;    // one work counter per bucket + [63] error flag
    // scratch
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    void *d_sort_ws = nullptr;
    size_t sort_ws_bytes = 0;
    double *h_pinned = nullptr;   // small pinned read-back buffer
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // bucket launches are spread over the main stream + aux streams so that their tails overlap
    static constexpr int kAux = 3;
    cudaStream_t aux[kAux] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[kAux] = {nullptr, nullptr, nullptr};
    int n_streams = 1;
    bool estep_timed = false, mstep_timed = false, elbo_valid = false;
    tmvb_stats st{};
};

namespace {

constexpr int kMaxBuckets = 48;

int ensure_scratch(tmvb_lda_t h, size_t bytes)
{
    if (bytes <= h->scratch_bytes) return 0;
    if (h->d_scratch) TMVB_CUDA(cudaFree(h->d_scratch));
    h->d_scratch = nullptr;
    h->scratch_bytes = 0;
    TMVB_CUDA(cudaMalloc(&h->d_scratch, bytes));
    h->scratch_bytes = bytes;
    return 0;
}

int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

LdaDev dev_view(tmvb_lda_t h)
{
    LdaDev p;
    p.K = (int)h->K;
    p.K_ld = h->K_ld;
    p.RS = h->RS;
    p.V = (int)h->V;
    p.M = h->M;
    p.beta = h->d_beta[h->cur];
    p.alpha = h->d_alpha;
    p.stats = h->d_stats;
    p.doc_off = h->d_doc_off;
    p.terms = h->d_terms;
    p.counts = h->d_counts;
    p.Elogtheta = h->d_Elogtheta;
    p.Elogtheta_old = h->d_Elogtheta_old;
    p.gamma = h->d_gamma;
    p.small = h->d_small;
    p.viter = 0;
    p.vtol = 0.f;
    p.stage_bulk = env_int("TMVB_LDA_STAGE_BULK", 1);
    p.dbg = env_int("TMVB_LDA_DBG", 0);
    return p;
}

int grid_for(long long work, int block, int n_sm)
{
    long long g = (work + block - 1) / block;
    long long cap = (long long)n_sm * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// Split the length-sorted documents into launches whose shared-memory tile capacity ("cap", in
// tokens) fits the longest document of the launch; more warps per CTA for longer documents.
int plan_buckets(tmvb_lda_t h, const std::vector<int> &len_sorted)
{
    h->buckets.clear();
    const int M = (int)len_sorted.size();
    if (M == 0) return 0;
    size_t budget = h->smem_optin;
    int cap_max = 16;
    while (lda_smem_bytes(h->RS, h->layout->lpt, cap_max + 16) <= budget) cap_max += 16;
    std::vector<int> caps;
    for (int c = 16; c < cap_max; c = (c < 128) ? c + 16 : (c < 256 ? c + 32 : c + c / 4 / 16 * 16)) caps.push_back(c);
    caps.push_back(cap_max);
    int begin = 0;  // documents are sorted by length, longest first
    for (int ci = (int)caps.size() - 1; ci >= 0 && begin < M; ci--) {
        const int lo = (ci == 0) ? -1 : caps[ci - 1];  // this launch takes lengths in (lo, caps[ci]] (+ overflow for the largest)
        int end = begin;
        while (end < M && len_sorted[end] > lo) end++;
        if (end == begin) continue;
        Bucket b;
        b.doc_begin = begin;
        b.doc_end = end;
        b.cap = std::min(caps[ci], std::max(16, (len_sorted[begin] + 15) / 16 * 16));
        if (b.cap > cap_max) b.cap = cap_max;
        b.warps = 1;
        b.smem = lda_smem_bytes(h->RS, h->layout->lpt, b.cap);
        b.grid = 0;
        h->buckets.push_back(b);
        begin = end;
    }
    if ((int)h->buckets.size() > kMaxBuckets) return fail(-1, "internal: too many launch buckets");
    return 0;
}

int free_all(tmvb_lda_t h)
{
    cudaFree(h->d_doc_off);
    cudaFree(h->d_src_off);
    cudaFree(h->d_terms);
    cudaFree(h->d_perm);
    cudaFree(h->d_counts);
    cudaFree(h->d_alpha);
    cudaFree(h->d_beta[0]);
    cudaFree(h->d_beta[1]);
    cudaFree(h->d_stats);
    cudaFree(h->d_Elogtheta);
    cudaFree(h->d_Elogtheta_old);
    cudaFree(h->d_gamma);
    cudaFree(h->d_small);
    cudaFree(h->d_local);
    cudaFree(h->d_counters);
    cudaFree(h->d_scratch);
    cudaFree(h->d_sort_ws);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    for (auto &e : h->ev)
        if (e) cudaEventDestroy(e);
    for (auto &a : h->aux)
        if (a) cudaStreamDestroy(a);
    for (auto &e : h->ev_join)
        if (e) cudaEventDestroy(e);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    return 0;
}

double lg_alpha_term(const std::vector<double> &a)
{
    double a0 = 0.0, sl = 0.0;
    for (double v : a) {
        a0 += v;
        sl += lgamma(v);
    }
    return lgamma(a0) - sl;  // LDA.jl:51
}

}  // namespace

extern "C" {

int tmvb_lda_create(tmvb_lda_t *out, int64_t K, int64_t M, int64_t V, int device, void *stream)
{
    TMVB_CHECK_ARG(out != nullptr, "handle pointer is NULL");
    *out = nullptr;
    TMVB_CHECK_ARG(K > 0, "number of topics must be a positive integer");  // gpuLDA.jl:47
    TMVB_CHECK_ARG(M >= 0 && V >= 0, "M and V must be nonnegative");
    TMVB_CHECK_ARG(M < (1ll << 31) && V < (1ll << 31), "M and V must fit in int32");
    if (K > 256) return fail(-2, "K=%lld is not supported (K <= 256)", (long long)K);
    const int K_ld = (int)((K + 7) / 8 * 8);
    int RS = 0;
    const Layout *layout = lda_pick_layout(K_ld, &RS);
    if (!layout) return fail(-2, "internal: no lane layout for K=%lld", (long long)K);
    int ndev = 0;
    TMVB_TRY(tmvb_device_count(&ndev));
    if (ndev == 0) return fail(-3, "no CUDA device: libtmvb has no CPU fallback");
    if (device < 0) TMVB_CUDA(cudaGetDevice(&device));
    TMVB_CHECK_ARG(device < ndev, "device index out of range");
    TMVB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TMVB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(-3, "device %d is sm_%d%d; libtmvb is built for sm_100a only", device, prop.major, prop.minor);

    tmvb_lda_t h = new tmvb_lda_s();
    h->device = device;
    h->n_sm = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    h->K = K;
    h->M = M;
    h->V = V;
    h->layout = layout;
    h->RS = RS;
    h->K_ld = K_ld;
    if (stream) {
        h->stream = (cudaStream_t)stream;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete h;
            return fail((int)e, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        h->own_stream = true;
    }
    const size_t kv = (size_t)std::max<int64_t>(V, 1) * h->K_ld, km = (size_t)std::max<int64_t>(M, 1) * h->K_ld;
    cudaError_t e = cudaSuccess;
    auto A = [&](void **p, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes, h->stream);
    };
    A((void **)&h->d_alpha, h->K_ld * 4);
    A((void **)&h->d_beta[0], kv * 4);
    A((void **)&h->d_beta[1], kv * 4);
    A((void **)&h->d_stats, kv * 4);
    A((void **)&h->d_Elogtheta, km * 4);
    A((void **)&h->d_Elogtheta_old, km * 4);
    A((void **)&h->d_gamma, km * 4);
    A((void **)&h->d_small, (h->K_ld + 2) * 8);
    A((void **)&h->d_local, (3 * h->K_ld + 2) * 8);
    A((void **)&h->d_counters, 64 * 4);
    if (e == cudaSuccess) e = cudaMallocHost((void **)&h->h_pinned, (3 * h->K_ld + 8) * 8);
    for (auto &evx : h->ev)
        if (e == cudaSuccess) e = cudaEventCreate(&evx);
    h->n_streams = std::min(1 + tmvb_lda_s::kAux, std::max(1, env_int("TMVB_LDA_STREAMS", 4)));
    for (int a = 0; a + 1 < h->n_streams; a++) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->aux[a], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join[a], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        free_all(h);
        delete h;
        return fail((int)e, "device allocation failed: %s", cudaGetErrorString(e));
    }
    h->h_alpha.assign(K, 1.0);
    // opt in to the large dynamic shared memory for both instantiations of this K
    for (int eb = 0; eb < 2; eb++) {
        EstepFn fn = layout->fn[eb];
        e = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin);
        if (e != cudaSuccess) {
            free_all(h);
            delete h;
            return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        }
    }
    *out = h;
    return 0;
}

int tmvb_lda_destroy(tmvb_lda_t h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_all(h);
    delete h;
    return 0;
}

int tmvb_lda_kld(tmvb_lda_t h, int64_t *K_ld)
{
    TMVB_CHECK_ARG(h && K_ld, "NULL argument");
    *K_ld = h->K_ld;
    return 0;
}

int tmvb_lda_set_corpus(tmvb_lda_t h, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CHECK_ARG(N_cumsum != nullptr, "N_cumsum is NULL");
    TMVB_CUDA(cudaSetDevice(h->device));
    const int64_t M = h->M;
    TMVB_CHECK_ARG(N_cumsum[0] == 0, "N_cumsum[0] must be 0");
    const int64_t nnz = N_cumsum[M];
    TMVB_CHECK_ARG(nnz >= 0, "N_cumsum must be nondecreasing");
    TMVB_CHECK_ARG(nnz == 0 || (terms != nullptr && counts != nullptr), "terms/counts are NULL");

    // host: O(M) counting sort of the documents by length (descending, stable)
    std::vector<int> len(M);
    int maxlen = 0;
    for (int64_t d = 0; d < M; d++) {
        const int64_t l = N_cumsum[d + 1] - N_cumsum[d];
        if (l < 0 || l > (1 << 30)) return fail(-1, "invalid argument: N_cumsum must be nondecreasing (document %lld)", (long long)d);
        len[d] = (int)l;
        maxlen = std::max(maxlen, (int)l);
    }
    std::vector<int64_t> start((size_t)maxlen + 2, 0);
    for (int64_t d = 0; d < M; d++) start[maxlen - len[d] + 1]++;
    for (int l = 0; l <= maxlen; l++) start[l + 1] += start[l];
    h->h_perm.assign(M, 0);
    for (int64_t d = 0; d < M; d++) h->h_perm[start[maxlen - len[d]]++] = (int)d;
    std::vector<long long> src_off(std::max<int64_t>(M, 1)), dst_off(M + 1);
    std::vector<int> len_sorted(M);
    dst_off[0] = 0;
    for (int64_t p = 0; p < M; p++) {
        const int d = h->h_perm[p];
        src_off[p] = N_cumsum[d];
        len_sorted[p] = len[d];
        dst_off[p + 1] = dst_off[p] + len[d];
    }
    TMVB_TRY(plan_buckets(h, len_sorted));

    const size_t nz = (size_t)std::max<int64_t>(nnz, 1);
    if (!h->d_doc_off) {
        TMVB_CUDA(cudaMalloc((void **)&h->d_doc_off, (M + 1) * 8));
        TMVB_CUDA(cudaMalloc((void **)&h->d_src_off, std::max<int64_t>(M, 1) * 8));
        TMVB_CUDA(cudaMalloc((void **)&h->d_perm, std::max<int64_t>(M, 1) * 4));
    }
    if (nz > h->nnz_cap) {  // token arrays are reused across calls while they fit
        cudaFree(h->d_terms);
        cudaFree(h->d_counts);
        h->d_terms = nullptr;
        h->d_counts = nullptr;
        h->nnz_cap = 0;
        TMVB_CUDA(cudaMalloc((void **)&h->d_terms, nz * 4));
        TMVB_CUDA(cudaMalloc((void **)&h->d_counts, nz * 4));
        h->nnz_cap = nz;
    }
    TMVB_CUDA(cudaMemcpyAsync(h->d_doc_off, dst_off.data(), (M + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    if (M > 0) {
        TMVB_CUDA(cudaMemcpyAsync(h->d_src_off, src_off.data(), M * 8, cudaMemcpyHostToDevice, h->stream));
        TMVB_CUDA(cudaMemcpyAsync(h->d_perm, h->h_perm.data(), M * 4, cudaMemcpyHostToDevice, h->stream));
    }
    h->st.h2d_bytes += (M + 1) * 8 + M * 12;
    if (nnz > 0) {
        TMVB_TRY(ensure_scratch(h, (size_t)nnz * 16));
        long long *t64 = (long long *)h->d_scratch, *c64 = t64 + nnz;
        TMVB_CUDA(cudaMemcpyAsync(t64, terms, nnz * 8, cudaMemcpyHostToDevice, h->stream));
        TMVB_CUDA(cudaMemcpyAsync(c64, counts, nnz * 8, cudaMemcpyHostToDevice, h->stream));
        h->st.h2d_bytes += nnz * 16;
        TMVB_CUDA(cudaMemsetAsync(h->d_counters + 63, 0, 4, h->stream));
        pack_corpus_kernel<<<grid_for(M * 32, 256, h->n_sm), 256, 0, h->stream>>>(t64, c64, h->d_src_off, h->d_doc_off, M, (int)h->V,
                                                                                 h->d_terms, h->d_counts, h->d_counters + 63);
        h->st.kernel_launches++;
        TMVB_CUDA(cudaGetLastError());
        int err = 0;
        TMVB_CUDA(cudaMemcpyAsync(&err, h->d_counters + 63, 4, cudaMemcpyDeviceToHost, h->stream));
        TMVB_CUDA(cudaStreamSynchronize(h->stream));  // also keeps the host vectors alive long enough
        if (err & 1) return fail(-1, "invalid argument: terms must lie in [0, V)");
        if (err & 2) return fail(-1, "invalid argument: all counts must be positive integers");  // Corpus.jl:43
    } else {
        TMVB_CUDA(cudaStreamSynchronize(h->stream));
    }
    h->nnz = nnz;
    h->corpus_set = true;
    return 0;
}

int tmvb_lda_set_alpha(tmvb_lda_t h, const float *alpha)
{
    TMVB_CHECK_ARG(h && alpha, "NULL argument");
    TMVB_CUDA(cudaSetDevice(h->device));
    for (int64_t i = 0; i < h->K; i++) {
        TMVB_CHECK_ARG(alpha[i] > 0.f && isfinite(alpha[i]), "alpha must be positive and finite");  // modelutils.jl:262-263
        h->h_alpha[i] = (double)alpha[i];
    }
    std::vector<float> pad(h->K_ld, 0.f);
    memcpy(pad.data(), alpha, h->K * 4);
    TMVB_CUDA(cudaMemcpyAsync(h->d_alpha, pad.data(), h->K_ld * 4, cudaMemcpyHostToDevice, h->stream));
    TMVB_CUDA(cudaStreamSynchronize(h->stream));
    h->st.h2d_bytes += h->K * 4;
    return 0;
}

int tmvb_lda_upload(tmvb_lda_t h, const float *alpha, const float *beta, const float *Elogtheta, const float *gamma)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CUDA(cudaSetDevice(h->device));
    if (alpha) TMVB_TRY(tmvb_lda_set_alpha(h, alpha));
    const int64_t K = h->K, M = h->M, V = h->V;
    if (beta && V > 0) {
        TMVB_TRY(ensure_scratch(h, (size_t)K * V * 4));
        TMVB_CUDA(cudaMemcpyAsync(h->d_scratch, beta, (size_t)K * V * 4, cudaMemcpyHostToDevice, h->stream));
        validate_kernel<<<grid_for(K * V, 256, h->n_sm), 256, 0, h->stream>>>((const float *)h->d_scratch, K * V, 0, h->d_counters + 62);
        pad_rows_kernel<<<grid_for(V * h->K_ld, 256, h->n_sm), 256, 0, h->stream>>>((const float *)h->d_scratch, h->d_beta[h->cur], nullptr, V, (int)K, h->K_ld);
        TMVB_CUDA(cudaGetLastError());
        // beta_old = copy(beta)  (LDA.jl:36)
        TMVB_CUDA(cudaMemcpyAsync(h->d_beta[h->cur ^ 1], h->d_beta[h->cur], (size_t)V * h->K_ld * 4, cudaMemcpyDeviceToDevice, h->stream));
        h->st.kernel_launches++;
        h->st.h2d_bytes += K * V * 4;
    }
    if ((Elogtheta || gamma) && M > 0) {
        TMVB_CHECK_ARG(h->corpus_set, "set_corpus must precede the upload of per-document parameters");
        TMVB_CUDA(cudaStreamSynchronize(h->stream));
        TMVB_TRY(ensure_scratch(h, (size_t)K * M * 4));
        if (Elogtheta) {
            TMVB_CUDA(cudaMemcpyAsync(h->d_scratch, Elogtheta, (size_t)K * M * 4, cudaMemcpyHostToDevice, h->stream));
            validate_kernel<<<grid_for(K * M, 256, h->n_sm), 256, 0, h->stream>>>((const float *)h->d_scratch, K * M, 1, h->d_counters + 62);
            pad_rows_kernel<<<grid_for(M * h->K_ld, 256, h->n_sm), 256, 0, h->stream>>>((const float *)h->d_scratch, h->d_Elogtheta, h->d_perm, M, (int)K, h->K_ld);
            TMVB_CUDA(cudaGetLastError());
            // Elogtheta_old = deepcopy(Elogtheta)  (LDA.jl:39)
            TMVB_CUDA(cudaMemcpyAsync(h->d_Elogtheta_old, h->d_Elogtheta, (size_t)M * h->K_ld * 4, cudaMemcpyDeviceToDevice, h->stream));
            h->st.kernel_launches++;
            h->st.h2d_bytes += K * M * 4;
        }
        if (gamma) {
            TMVB_CUDA(cudaMemcpyAsync(h->d_scratch, gamma, (size_t)K * M * 4, cudaMemcpyHostToDevice, h->stream));
            validate_kernel<<<grid_for(K * M, 256, h->n_sm), 256, 0, h->stream>>>((const float *)h->d_scratch, K * M, 2, h->d_counters + 62);
            pad_rows_kernel<<<grid_for(M * h->K_ld, 256, h->n_sm), 256, 0, h->stream>>>((const float *)h->d_scratch, h->d_gamma, h->d_perm, M, (int)K, h->K_ld);
            TMVB_CUDA(cudaGetLastError());
            h->st.kernel_launches++;
            h->st.h2d_bytes += K * M * 4;
        }
    }
    int verr = 0;
    TMVB_CUDA(cudaMemcpyAsync(&verr, h->d_counters + 62, 4, cudaMemcpyDeviceToHost, h->stream));
    TMVB_CUDA(cudaStreamSynchronize(h->stream));
    TMVB_CUDA(cudaMemsetAsync(h->d_counters + 62, 0, 4, h->stream));
    // the messages of check_model(::gpuLDA), modelutils.jl:264-273
    if (verr & 0x3) return fail(-5, "beta must be a right stochastic matrix.");
    if (verr & 0x4) return fail(-5, "Elogtheta must be finite.");
    if (verr & 0x8) return fail(-5, "Elogtheta must be nonpositive.");
    if (verr & 0x10) return fail(-5, "gamma must be finite.");
    if (verr & 0x20) return fail(-5, "gamma must be positive.");
    h->params_set = true;
    h->elbo_valid = false;
    return 0;
}

int tmvb_lda_estep(tmvb_lda_t h, int viter, float vtol, int want_elbo)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CHECK_ARG(viter >= 1, "viter must be at least 1");
    TMVB_CHECK_ARG(vtol >= 0.f, "tolerance parameters must be nonnegative");  // gpuLDA.jl:349
    TMVB_CHECK_ARG(h->corpus_set, "set_corpus has not been called");
    TMVB_CUDA(cudaSetDevice(h->device));
    LdaDev p = dev_view(h);
    p.viter = viter;
    p.vtol = vtol;
    TMVB_CUDA(cudaEventRecord(h->ev[0], h->stream));
    TMVB_CUDA(cudaMemsetAsync(h->d_small, 0, (h->K_ld + 2) * 8, h->stream));
    TMVB_CUDA(cudaMemsetAsync(h->d_counters, 0, 64 * 4, h->stream));
    h->h_alpha_estep = h->h_alpha;
    EstepFn fn = h->layout->fn[want_elbo != 0];
    const int ns = (h->buckets.size() > 1) ? h->n_streams : 1;
    if (ns > 1) {
        TMVB_CUDA(cudaEventRecord(h->ev_fork, h->stream));
        for (int a = 0; a + 1 < ns; a++) TMVB_CUDA(cudaStreamWaitEvent(h->aux[a], h->ev_fork, 0));
    }
    for (size_t bi = 0; bi < h->buckets.size(); bi++) {
        Bucket &b = h->buckets[bi];
        const int threads = 32 * b.warps;
        if (b.grid == 0) {
            int occ = 0;
            TMVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, threads, b.smem));
            if (occ < 1) return fail(-4, "internal: E-step kernel does not fit (cap=%d warps=%d smem=%zu)", b.cap, b.warps, b.smem);
            b.grid = std::min(b.doc_end - b.doc_begin, occ * h->n_sm);
        }
        void *args[] = {(void *)&p, (void *)&b.doc_begin, (void *)&b.doc_end, (void *)&b.cap, (void *)nullptr};
        int *counter = h->d_counters + bi;
        args[4] = (void *)&counter;
        cudaStream_t st = (bi % ns == 0) ? h->stream : h->aux[bi % ns - 1];
        TMVB_CUDA(cudaLaunchKernel((const void *)fn, dim3(b.grid), dim3(threads), args, b.smem, st));
        h->st.kernel_launches++;
    }
    for (int a = 0; a + 1 < ns; a++) {
        TMVB_CUDA(cudaEventRecord(h->ev_join[a], h->aux[a]));
        TMVB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join[a], 0));
    }
    TMVB_CUDA(cudaEventRecord(h->ev[1], h->stream));
    h->estep_timed = true;
    h->elbo_valid = (want_elbo != 0);
    return 0;
}

int tmvb_lda_reduce_buffers(tmvb_lda_t h, void **stats, int64_t *n_stats, void **small, int64_t *n_small)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    if (stats) *stats = h->d_stats;
    if (n_stats) *n_stats = (int64_t)h->V * h->K_ld;
    if (small) *small = h->d_small;
    if (n_small) *n_small = h->K_ld + 2;
    return 0;
}

int tmvb_lda_mstep(tmvb_lda_t h)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CUDA(cudaSetDevice(h->device));
    TMVB_CUDA(cudaEventRecord(h->ev[2], h->stream));
    if (h->V > 0) {
        TMVB_CUDA(cudaMemsetAsync(h->d_local, 0, (2 * h->K_ld) * 8, h->stream));
        const int R = std::max(1, 256 / h->K_ld);
        const int threads = std::max(R * h->K_ld, h->K_ld);
        const int grid = (int)std::min<int64_t>((h->V + R - 1) / R, (int64_t)h->n_sm * 8);
        lda_colsum_kernel<<<grid, threads, threads * 8, h->stream>>>(h->d_stats, (int)h->V, h->K_ld, h->d_local);
        TMVB_CUDA(cudaGetLastError());
        const long long n = (long long)h->V * h->K_ld;
        lda_normalize_kernel<<<grid_for(n, 256, h->n_sm), 256, 0, h->stream>>>(h->d_stats, h->d_beta[h->cur ^ 1], h->d_local, n, (int)h->K, h->K_ld,
                                                                              h->d_local + h->K_ld, h->elbo_valid ? 1 : 0);
        TMVB_CUDA(cudaGetLastError());
        h->st.kernel_launches += 2;
        h->cur ^= 1;  // beta_old <- beta ; beta <- new  (LDA.jl:122-123)
    }
    TMVB_CUDA(cudaEventRecord(h->ev[3], h->stream));
    h->mstep_timed = true;
    return 0;
}

int tmvb_lda_get_elogtheta_sum(tmvb_lda_t h, double *out)
{
    TMVB_CHECK_ARG(h && out, "NULL argument");
    TMVB_CUDA(cudaSetDevice(h->device));
    TMVB_CUDA(cudaMemcpyAsync(h->h_pinned, h->d_small, (h->K_ld + 2) * 8, cudaMemcpyDeviceToHost, h->stream));
    TMVB_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(out, h->h_pinned, h->K * 8);
    h->st.d2h_bytes += (h->K_ld + 2) * 8;
    return 0;
}

int tmvb_lda_update_alpha(tmvb_lda_t h, int64_t M_total, int niter, double ntol, float *alpha_out)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CHECK_ARG(niter >= 0 && ntol >= 0.0, "iteration/tolerance parameters must be nonnegative");
    const int K = (int)h->K;
    std::vector<double> Esum(K), grad(K), hinv(K), pdir(K);
    TMVB_TRY(tmvb_lda_get_elogtheta_sum(h, Esum.data()));
    std::vector<double> &a = h->h_alpha;
    // LDA.jl:97-118: interior-point Newton, log barrier nu halved every step
    double nu = (double)K;
    const double Md = (double)M_total;
    for (int it = 0; it < niter; it++) {
        double rho = 1.0, a0 = 0.0;
        for (int i = 0; i < K; i++) a0 += a[i];
        const double dg0 = h_digamma(a0);
        double gh = 0.0, hs = 0.0, gn = 0.0;
        for (int i = 0; i < K; i++) {
            grad[i] = nu / a[i] + Md * (dg0 - h_digamma(a[i])) + Esum[i];
            hinv[i] = -1.0 / (Md * h_trigamma(a[i]) + nu / (a[i] * a[i]));
            gh += grad[i] * hinv[i];
            hs += hinv[i];
            gn += grad[i] * grad[i];
        }
        const double z = gh / (1.0 / (Md * h_trigamma(a0)) + hs);
        for (int i = 0; i < K; i++) pdir[i] = (grad[i] - z) * hinv[i];
        for (;;) {
            double mn = INFINITY;
            for (int i = 0; i < K; i++) mn = std::min(mn, a[i] - rho * pdir[i]);
            if (!(mn < 0.0)) break;
            rho *= 0.5;
        }
        for (int i = 0; i < K; i++) a[i] = copysign(std::min(fabs(a[i] - rho * pdir[i]), 1.7976931348623157e308), a[i]);
        if ((rho * sqrt(gn) < ntol) && (nu / (double)K < ntol)) break;
        nu *= 0.5;
    }
    for (int i = 0; i < K; i++) a[i] += TMVB_EPS_D;
    std::vector<float> pad(h->K_ld, 0.f);
    for (int i = 0; i < K; i++) {
        // keep the fp32 device copy strictly positive (the fp64 iterate can sit below FLT_MIN)
        pad[i] = std::max((float)a[i], 1.1754944e-38f);
        if (alpha_out) alpha_out[i] = pad[i];
    }
    TMVB_CUDA(cudaMemcpyAsync(h->d_alpha, pad.data(), h->K_ld * 4, cudaMemcpyHostToDevice, h->stream));
    TMVB_CUDA(cudaStreamSynchronize(h->stream));
    h->st.h2d_bytes += h->K * 4;
    return 0;
}

int tmvb_lda_elbo(tmvb_lda_t h, int mode, int64_t M_total, double *elbo_docs, double *elbo_global)
{
    TMVB_CHECK_ARG(h && elbo_docs && elbo_global, "NULL argument");
    TMVB_CHECK_ARG(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
    TMVB_CUDA(cudaSetDevice(h->device));
    const int K = (int)h->K, K_ld = h->K_ld;
    if (mode == 0) {
        TMVB_CHECK_ARG(h->elbo_valid, "mode 0 needs estep(want_elbo=1) followed by mstep");
        TMVB_CUDA(cudaMemcpyAsync(h->h_pinned, h->d_small, (K_ld + 2) * 8, cudaMemcpyDeviceToHost, h->stream));
        TMVB_CUDA(cudaMemcpyAsync(h->h_pinned + K_ld + 2, h->d_local + K_ld, 8, cudaMemcpyDeviceToHost, h->stream));
        TMVB_CUDA(cudaStreamSynchronize(h->stream));
        h->st.d2h_bytes += (K_ld + 3) * 8;
        const double *Esum = h->h_pinned;
        double g = (double)M_total * lg_alpha_term(h->h_alpha);           // Elogptheta, LDA.jl:51
        // dot(alpha .- 1, Elogtheta[d]) summed over d, plus the (1 - alpha_estep) . Elogtheta_sum left
        // over from the per-document entropy/Elogpz terms (see lda_estep_kernel)
        for (int i = 0; i < K; i++) g += (h->h_alpha[i] - h->h_alpha_estep[i] - TMVB_EPS_D) * Esum[i];
        g += h->h_pinned[K_ld + 2];                                       // Elogpw over the statistics
        *elbo_docs = h->h_pinned[K_ld];
        *elbo_global = g;
        return 0;
    }
    TMVB_CHECK_ARG(h->corpus_set, "set_corpus has not been called");
    LdaDev p = dev_view(h);
    double *out = h->d_local + 2 * K_ld;
    TMVB_CUDA(cudaMemsetAsync(out, 0, 8, h->stream));
    if (h->M > 0) {
        if (mode == 2)
            lda_elbo_kernel<double><<<grid_for(h->M * 32, 128, h->n_sm), 128, 0, h->stream>>>(p, h->d_beta[h->cur ^ 1], lg_alpha_term(h->h_alpha), out);
        else
            lda_elbo_kernel<float><<<grid_for(h->M * 32, 128, h->n_sm), 128, 0, h->stream>>>(p, h->d_beta[h->cur ^ 1], lg_alpha_term(h->h_alpha), out);
        TMVB_CUDA(cudaGetLastError());
        h->st.kernel_launches++;
    }
    TMVB_CUDA(cudaMemcpyAsync(h->h_pinned, out, 8, cudaMemcpyDeviceToHost, h->stream));
    TMVB_CUDA(cudaStreamSynchronize(h->stream));
    h->st.d2h_bytes += 8;
    *elbo_docs = h->h_pinned[0];
    *elbo_global = 0.0;
    return 0;
}

static int download_rows(tmvb_lda_t h, const float *d_src, float *dst, int64_t rows, const int *perm)
{
    if (!dst || rows == 0) return 0;
    TMVB_TRY(ensure_scratch(h, (size_t)rows * h->K * 4));
    unpad_rows_kernel<<<grid_for(rows * h->K, 256, h->n_sm), 256, 0, h->stream>>>(d_src, (float *)h->d_scratch, perm, rows, (int)h->K, h->K_ld);
    TMVB_CUDA(cudaGetLastError());
    h->st.kernel_launches++;
    TMVB_CUDA(cudaMemcpyAsync(dst, h->d_scratch, (size_t)rows * h->K * 4, cudaMemcpyDeviceToHost, h->stream));
    TMVB_CUDA(cudaStreamSynchronize(h->stream));
    h->st.d2h_bytes += rows * h->K * 4;
    return 0;
}

int tmvb_lda_download(tmvb_lda_t h, float *alpha, float *beta, float *Elogtheta, float *gamma)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CUDA(cudaSetDevice(h->device));
    if (alpha) {
        TMVB_CUDA(cudaStreamSynchronize(h->stream));
        for (int64_t i = 0; i < h->K; i++) alpha[i] = std::max((float)h->h_alpha[i], 1.1754944e-38f);
    }
    TMVB_TRY(download_rows(h, h->d_beta[h->cur], beta, h->V, nullptr));
    if (h->M > 0 && (Elogtheta || gamma)) TMVB_CHECK_ARG(h->corpus_set, "set_corpus has not been called");
    TMVB_TRY(download_rows(h, h->d_Elogtheta, Elogtheta, h->M, h->d_perm));
    TMVB_TRY(download_rows(h, h->d_gamma, gamma, h->M, h->d_perm));
    return 0;
}

int tmvb_lda_download_old(tmvb_lda_t h, float *beta_old, float *Elogtheta_old)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CUDA(cudaSetDevice(h->device));
    TMVB_TRY(download_rows(h, h->d_beta[h->cur ^ 1], beta_old, h->V, nullptr));
    TMVB_TRY(download_rows(h, h->d_Elogtheta_old, Elogtheta_old, h->M, h->d_perm));
    return 0;
}

int tmvb_lda_materialize_phi(tmvb_lda_t h, float *phi)
{
    TMVB_CHECK_ARG(h && phi, "NULL argument");
    TMVB_CHECK_ARG(h->corpus_set, "set_corpus has not been called");
    TMVB_CUDA(cudaSetDevice(h->device));
    if (h->nnz == 0) return 0;
    const size_t bytes = (size_t)h->nnz * h->K * 4;
    TMVB_TRY(ensure_scratch(h, bytes));
    LdaDev p = dev_view(h);
    lda_phi_kernel<<<grid_for(h->M * 32, 128, h->n_sm), 128, 0, h->stream>>>(p, h->d_beta[h->cur ^ 1], h->d_src_off, (float *)h->d_scratch);
    TMVB_CUDA(cudaGetLastError());
    h->st.kernel_launches++;
    TMVB_CUDA(cudaMemcpyAsync(phi, h->d_scratch, bytes, cudaMemcpyDeviceToHost, h->stream));
    TMVB_CUDA(cudaStreamSynchronize(h->stream));
    h->st.d2h_bytes += bytes;
    return 0;
}

int tmvb_lda_topics(tmvb_lda_t h, int32_t *topics)
{
    TMVB_CHECK_ARG(h && topics, "NULL argument");
    TMVB_CUDA(cudaSetDevice(h->device));
    if (h->V == 0) return 0;
    const size_t bytes = (size_t)h->K * h->V * 4;
    TMVB_TRY(ensure_scratch(h, bytes));
    TMVB_TRY(topics_argsort(h->d_beta[h->cur], nullptr, (int)h->K, h->K_ld, (int)h->V, (int *)h->d_scratch, &h->d_sort_ws, &h->sort_ws_bytes,
                            h->stream, h->n_sm));
    h->st.kernel_launches += 3;
    TMVB_CUDA(cudaMemcpyAsync(topics, h->d_scratch, bytes, cudaMemcpyDeviceToHost, h->stream));
    TMVB_CUDA(cudaStreamSynchronize(h->stream));
    h->st.d2h_bytes += bytes;
    return 0;
}

int tmvb_lda_sync(tmvb_lda_t h)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CUDA(cudaSetDevice(h->device));
    TMVB_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int tmvb_lda_get_stats(tmvb_lda_t h, tmvb_stats *out)
{
    TMVB_CHECK_ARG(h && out, "NULL argument");
    TMVB_CUDA(cudaSetDevice(h->device));
    TMVB_CUDA(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    if (h->estep_timed) {
        TMVB_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        h->st.estep_ms = ms;
        TMVB_CUDA(cudaMemcpy(h->h_pinned, h->d_small + h->K_ld + 1, 8, cudaMemcpyDeviceToHost));
        h->st.sweeps = (int64_t)h->h_pinned[0];
    }
    if (h->mstep_timed) {
        TMVB_CUDA(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]));
        h->st.mstep_ms = ms;
    }
    *out = h->st;
    return 0;
}

}  // extern "C"
