// tmvb_lda.cu -- LDA coordinate-ascent VB on sm_100a: fused per-document E-step + statistic scatter, ELBO, and
// the tmvb_lda_* C ABI (include/tmvb.h).  Shared building blocks: tmvb_estep.cuh (token passes, lane layouts),
// tmvb_shard.cu (corpus re-layout, buckets, M-step normalisation).
//
// Reference semantics followed: the CPU model src/LDA.jl (per-document stopping rule, lagged-phi ELBO); the thing
// replaced: src/gpuLDA.jl's 7 OpenCL kernels + modelutils.jl:370-397,501-516.
//
// Device data layout (all owned by the handle):
//   beta[2]      float [V][K_ld]   term-major rows ("term rows", == Julia's column-major K x V with the leading
//                                  dimension padded to K_ld = 8*ceil(K/8), pad = 0); double buffered so beta_old
//                                  (LDA.jl:122) costs nothing
//   stats        float [V][K_ld]   sufficient statistics beta_temp (LDA.jl:131), REDG.ADD target
//   Elogtheta, Elogtheta_old, gamma  float [M][K_ld]   per-document K-vectors, internal document order
//   doc_off int64 [M+1], terms int32 [nnz], counts float [nnz]   CSR, documents sorted by length (descending)
//   small        double [K_ld+2]   sum_d Elogtheta_d | per-document ELBO terms | sweep counter
//                                  (summed across ranks together with stats in multi-GPU runs)
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "tmvb_lda_estep.cuh"
#include "tmvb_lda_hyb.cuh"

namespace tmvb {

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
template <typename real, int RM>
__global__ void lda_elbo_kernel(const LdaDev p, const float *__restrict__ beta_old, double lg_alpha_term, double *out)
{
    constexpr bool F64 = sizeof(real) == 8;  // RM = ceil(K / 32) rounded up to {1, 2, 4, 8}: topics i = lane + 32 r held in registers
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const real eps = (real)TMVB_EPS_D;
    double acc = 0.0;
    for (long long d = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); d < p.M; d += (long long)gridDim.x * wpb) {
        const long long o = p.doc_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o);
        const float *En = p.Elogtheta + d * p.K_ld, *Eo = p.Elogtheta_old + d * p.K_ld, *gm = p.gamma + d * p.K_ld;
        double dacc = 0.0, g0 = 0.0;
        real e_r[RM], En_r[RM];
#pragma unroll
        for (int r = 0; r < RM; r++) {
            const int i = lane + 32 * r;
            e_r[r] = 0;
            En_r[r] = 0;
            if (i < p.K) {
                const double g = gm[i], E = En[i];
                g0 += g;
                e_r[r] = F64 ? (real)exp((double)Eo[i]) : (real)expf(Eo[i]);
                En_r[r] = (real)E;
                if (F64) {
                    dacc += ((double)p.alpha[i] - 1.0) * E + (p.K > 1 ? lgamma(g) - (g - 1.0) * d_digamma(g) : 0.0);
                } else {
                    const PsiLg pl = psi_lgamma<true>((float)g);
                    dacc += ((double)p.alpha[i] - 1.0) * E + (p.K > 1 ? (double)pl.lg - (g - 1.0) * (double)pl.psi : 0.0);   // K = 1: entropy(Dirichlet) = 0, utils.jl:168
                }
            }
        }
        g0 = warp_sum_d(g0);
        real tacc = 0;
        // UN tokens per iteration: their row loads and warp reductions are independent, which hides the
        // load -> reduce -> log latency chain that made this pass slower than a whole E-step
        constexpr int UN = F64 ? 1 : 4;
        for (int n0 = 0; n0 < Nd; n0 += UN) {
            real u_r[UN][RM], s[UN], c[UN];
            const float *bn[UN];
#pragma unroll
            for (int q = 0; q < UN; q++) {
                const int n = min(n0 + q, Nd - 1);
                const int term = p.terms[o + n];
                c[q] = (n0 + q < Nd) ? (real)p.counts[o + n] : (real)0;
                const float *bo = beta_old + (size_t)term * p.K_ld;
                bn[q] = p.beta + (size_t)term * p.K_ld;
                s[q] = 0;
#pragma unroll
                for (int r = 0; r < RM; r++) {
                    const int i = lane + 32 * r;
                    u_r[q][r] = (i < p.K) ? eps + (real)bo[i] * e_r[r] : (real)0;
                    s[q] += u_r[q][r];
                }
            }
#pragma unroll
            for (int q = 0; q < UN; q++) s[q] = F64 ? (real)warp_sum_d((double)s[q]) : (real)warp_sum((float)s[q]);
#pragma unroll
            for (int q = 0; q < UN; q++) {
                // sum_i phi_i (E_i + ln(beta_i + eps) - ln phi_i),  phi = u / s,  ln phi = ln u - ln s
                const real ls = F64 ? (real)log((double)s[q]) : (real)__logf((float)s[q]);
                real a = 0;
#pragma unroll
                for (int r = 0; r < RM; r++) {
                    const int i = lane + 32 * r;
                    if (i < p.K) {
                        const real lb = F64 ? (real)log((double)bn[q][i] + TMVB_EPS_D) : (real)__logf(bn[q][i] + TMVB_EPS);
                        const real lu = F64 ? (real)log((double)u_r[q][r]) : (real)__logf((float)u_r[q][r]);
                        a += u_r[q][r] * (En_r[r] + lb - lu + ls);
                    }
                }
                if (F64) dacc += (double)(c[q] * a / s[q]); else tacc += c[q] * a / s[q];
            }
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

// psi(x) and psi'(x) together in fp64 without a division per recurrence step: with P(x) = x (x+1) ... (x+9),
//   psi(x) = psi(x+10) - P'/P ,   psi'(x) = psi'(x+10) + (P'^2 - P P'') / P^2 ,
// P, P', P'' by the product rule (fixed 10 steps, branch-free), then the Bernoulli series at y = x + 10 >= 10.
__device__ __forceinline__ void d_psi_tri(double x, double &psi, double &tri)
{
    double P = x, D1 = 1.0, D2 = 0.0;
#pragma unroll
    for (int k = 1; k < 10; k++) {
        const double f = x + (double)k;
        D2 = fma(D2, f, 2.0 * D1);
        D1 = fma(D1, f, P);
        P *= f;
    }
    const double y = x + 10.0, t = 1.0 / y, t2 = t * t, iP = 1.0 / P, q = D1 * iP;
    const double sp = t2 * (1.0 / 12 - t2 * (1.0 / 120 - t2 * (1.0 / 252 - t2 * (1.0 / 240 - t2 * (1.0 / 132 - t2 * (691.0 / 32760 - t2 * (1.0 / 12)))))));
    psi = log(y) - 0.5 * t - sp - q;
    const double st = t * (1.0 + 0.5 * t + t2 * (1.0 / 6 - t2 * (1.0 / 30 - t2 * (1.0 / 42 - t2 * (1.0 / 30 - t2 * (5.0 / 66 - t2 * (691.0 / 2730 - t2 * (7.0 / 6))))))));
    tri = st + (q * q - D2 * iP);
}
__device__ __forceinline__ double warp_min_d(double v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

// update_elbo! for an arbitrary device state in one cheap pass (tmvb_lda_elbo mode 1, the `update_elbo!` at the top of
// train!, gpuLDA.jl:353).  Same expectations as lda_elbo_kernel, with the two logarithms per (token, topic) moved into a
// K x V table:  ln(beta_i,w + eps) - ln u_ni = D[w][i] - Elogtheta_old_i,  D = ln(beta + eps) - ln(beta_old + eps),
// exact up to terms weighted by phi_ni ~ eps / s_n (u_ni = eps + beta_old_i,w e_i).  Per token: two FMA passes over the
// row and one logarithm.  lda_elbo_kernel<double> (mode 2) stays the literal restatement the tests compare against.
__global__ void lda_logratio_kernel(const float *__restrict__ beta, const float *__restrict__ beta_old, float *__restrict__ D, long long n,
                                    int K, int K_ld)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q % K_ld);
        D[q] = (i < K) ? logf(beta[q] + TMVB_EPS) - logf(beta_old[q] + TMVB_EPS) : 0.0f;
    }
}

template <int RM>
__global__ void lda_elbo_fast_kernel(const LdaDev p, const float *__restrict__ beta_old, const float *__restrict__ D, double lg_alpha_term,
                                     double *out)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    double acc = 0.0;
    for (long long d = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); d < p.M; d += (long long)gridDim.x * wpb) {
        const long long o = p.doc_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o);
        const float *En = p.Elogtheta + d * p.K_ld, *Eo = p.Elogtheta_old + d * p.K_ld, *gm = p.gamma + d * p.K_ld;
        double dacc = 0.0, g0 = 0.0;
        float e_r[RM], dE_r[RM];
#pragma unroll
        for (int r = 0; r < RM; r++) {
            const int i = lane + 32 * r;
            e_r[r] = dE_r[r] = 0.0f;
            if (i < p.K) {
                const double g = gm[i], E = En[i];
                g0 += g;
                e_r[r] = expf(Eo[i]);
                dE_r[r] = En[i] - Eo[i];
                const PsiLg pl = psi_lgamma<true>((float)g);
                dacc += ((double)p.alpha[i] - 1.0) * E + (p.K > 1 ? (double)pl.lg - (g - 1.0) * (double)pl.psi : 0.0);   // K = 1: entropy(Dirichlet) = 0, utils.jl:168
            }
        }
        g0 = warp_sum_d(g0);
        float tacc = 0.0f;
        constexpr int UN = 4;
        for (int n0 = 0; n0 < Nd; n0 += UN) {
            float s[UN], a[UN], c[UN];
#pragma unroll
            for (int q = 0; q < UN; q++) {
                const int n = min(n0 + q, Nd - 1);
                const int term = p.terms[o + n];
                c[q] = (n0 + q < Nd) ? p.counts[o + n] : 0.0f;
                const float *bo = beta_old + (size_t)term * p.K_ld, *dr = D + (size_t)term * p.K_ld;
                s[q] = a[q] = 0.0f;
#pragma unroll
                for (int r = 0; r < RM; r++) {
                    const int i = lane + 32 * r;
                    if (i < p.K) {
                        const float u = fmaf(bo[i], e_r[r], TMVB_EPS);
                        s[q] += u;
                        a[q] = fmaf(u, dE_r[r] + dr[i], a[q]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < UN; q++) {
                s[q] = warp_sum(s[q]);
                a[q] = warp_sum(a[q]);
            }
#pragma unroll
            for (int q = 0; q < UN; q++) tacc += c[q] * (__fdividef(a[q], s[q]) + __logf(s[q]));
        }
        if (lane == 0) dacc += (double)tacc;
        dacc = warp_sum_d(dacc);
        if (lane == 0) {
            double ent = 0.0;
            if (p.K > 1) ent = -lgamma(g0) + (g0 - (double)p.K) * d_digamma(g0);
            acc += dacc + ent + lg_alpha_term;
        }
    }
    if (lane == 0 && acc != 0.0) atomicAdd(out, acc);
}

// update_elbo! for a state whose lagged copies EQUAL the current ones (right after tmvb_lda_upload: beta_old = copy(beta),
// Elogtheta_old = deepcopy(Elogtheta), LDA.jl:36,39) -- the `update_elbo!` at the top of train! (gpuLDA.jl:353) in every
// train call.  With beta_old = beta and Elogtheta_old = Elogtheta the table D and dE of lda_elbo_fast_kernel vanish and the
// token terms collapse to  sum_n c_n ln s_n,  s_n = K eps + sum_i beta_i,w exp(Elogtheta_i)  (same eps-weighted remainder as
// there): one dot product per token in the E-step's lane layout (LPT lanes x CPL 16-byte chunks per row, rows straight from
// L2), UN rounds in flight.  The per-document Dirichlet terms are those of lda_elbo_fast_kernel.
template <int LPT, int CPL>
__global__ void __launch_bounds__(128) lda_elbo_fresh_kernel(const LdaDev p, double lg_alpha_term, double *out)
{
    constexpr int S = 32 / LPT, UN = 4;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int kl = lane % LPT, ts = lane / LPT;
    const int K = p.K, K_ld = p.K_ld, CH = K_ld >> 2;
    const float Keps = (float)K * TMVB_EPS;
    const ulonglong2 zero = make_ulonglong2(0ull, 0ull);
    double acc = 0.0;
    for (long long d = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); d < p.M; d += (long long)gridDim.x * wpb) {
        const long long o = p.doc_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o);
        const float *En = p.Elogtheta + d * K_ld, *gm = p.gamma + d * K_ld;
        double dacc = 0.0, g0 = 0.0;
        for (int i = lane; i < K; i += 32) {
            const double g = gm[i], E = En[i];
            g0 += g;
            const PsiLg pl = psi_lgamma<true>((float)g);
            dacc += ((double)p.alpha[i] - 1.0) * E + (p.K > 1 ? (double)pl.lg - (g - 1.0) * (double)pl.psi : 0.0);   // K = 1: entropy(Dirichlet) = 0, utils.jl:168
        }
        g0 = warp_sum_d(g0);
        f32x2 e01[CPL], e23[CPL];
#pragma unroll
        for (int m = 0; m < CPL; m++) {
            const int q = kl + LPT * m;
            float4 E = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool in = (m < CPL - 1 || q < CH);
            if (in) E = *reinterpret_cast<const float4 *>(En + 4 * q);
            // pad topics (i >= K) have beta = 0: their e does not matter
            e01[m] = in ? pk2(__expf(E.x), __expf(E.y)) : 0ull;
            e23[m] = in ? pk2(__expf(E.z), __expf(E.w)) : 0ull;
        }
        float tacc = 0.0f;
        const int rounds = (Nd + S - 1) / S;
        for (int r0 = 0; r0 < rounds; r0 += UN) {
            ulonglong2 b[UN][CPL];
            float c[UN];
#pragma unroll
            for (int u = 0; u < UN; u++) {
                const int n = (r0 + u) * S + ts;
                const bool ok = n < Nd;
                const int term = ok ? __ldg(p.terms + o + n) : 0;
                c[u] = ok ? __ldg(p.counts + o + n) : 0.0f;
                const ulonglong2 *row = reinterpret_cast<const ulonglong2 *>(p.beta + (size_t)term * K_ld) + kl;
#pragma unroll
                for (int m = 0; m < CPL; m++) b[u][m] = (ok && (m < CPL - 1 || kl + LPT * m < CH)) ? __ldg(row + LPT * m) : zero;
            }
#pragma unroll
            for (int u = 0; u < UN; u++) {
                const float sn = tok_dot<LPT, CPL>(b[u], e01, e23) + Keps;
                if (kl == 0 && c[u] > 0.0f) tacc = fmaf(c[u], __logf(sn), tacc);
            }
        }
        dacc += (double)tacc;
        dacc = warp_sum_d(dacc);
        if (lane == 0) {
            double ent = 0.0;
            if (K > 1) ent = -lgamma(g0) + (g0 - (double)K) * d_digamma(g0);
            acc += dacc + ent + lg_alpha_term;
        }
    }
    if (lane == 0 && acc != 0.0) atomicAdd(out, acc);
}
typedef void (*LdaElboFreshFn)(const LdaDev, double, double *);
#define TMVB_LDA_FRESH_FN(L, C) (LdaElboFreshFn)lda_elbo_fresh_kernel<L, C>,
static const LdaElboFreshFn kLdaElboFresh[kNumLaneLayouts] = {TMVB_FOR_EACH_LAYOUT(TMVB_LDA_FRESH_FN)};

// Block-wide fp64 reductions for lda_alpha_kernel (blockDim.x = 32 * nw, nw <= 9): warp shuffles, then one shared-memory
// round; every thread receives the result.  `red` is double[3][16], two barriers per call.
__device__ __forceinline__ void block_sum3(double &a, double &b, double &c, double (*red)[16])
{
    a = warp_sum_d(a);
    b = warp_sum_d(b);
    c = warp_sum_d(c);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (nw == 1) return;
    __syncthreads();
    if (lane == 0) {
        red[0][warp] = a;
        red[1][warp] = b;
        red[2][warp] = c;
    }
    __syncthreads();
    a = b = c = 0.0;
    for (int w = 0; w < nw; w++) {
        a += red[0][w];
        b += red[1][w];
        c += red[2][w];
    }
}
__device__ __forceinline__ double block_min(double v, double (*red)[16])
{
    v = warp_min_d(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (nw == 1) return v;
    __syncthreads();
    if (lane == 0) red[0][warp] = v;
    __syncthreads();
    v = red[0][0];
    for (int w = 1; w < nw; w++) v = fmin(v, red[0][w]);
    return v;
}

// update_alpha! (LDA.jl:97-118): interior-point Newton with log barrier in fp64.  One CTA: thread i < K owns alpha_i, thread K
// owns sum(alpha), so that every Newton step costs ONE digamma/trigamma evaluation per thread (the one-warp version
// spent 57 us per outer iteration on three serial evaluations per lane).  Returns (to every thread) the part of the ELBO
// that depends on alpha:  M (lnG(sum alpha) - sum lnG(alpha)) + sum_i (alpha_i - alpha_estep_i - eps) Esum_i   (see tmvb_lda_elbo).
//   Es = this thread's sum_d Elogtheta_di (i = threadIdx.x < K)
__device__ __forceinline__ double lda_alpha_newton(double *__restrict__ alpha64, float *__restrict__ alpha32, double Es, int K, double Md, int niter,
                                                   double ntol, double (*red)[16], double *bc)
{
    const int tid = threadIdx.x;
    const bool is_topic = tid < K, is_sum = tid == K;
    double a = is_topic ? alpha64[tid] : 1.0;
    const double a_estep = a;
    double nu = (double)K;
    for (int it = 0; it < niter; it++) {
        double a0 = is_topic ? a : 0.0, u1 = 0.0, u2 = 0.0;
        block_sum3(a0, u1, u2, red);
        double dg = 0.0, tg = 1.0;
        if (is_topic || is_sum) d_psi_tri(is_topic ? a : a0, dg, tg);
        __syncthreads();
        if (is_sum) {
            bc[0] = dg;
            bc[1] = tg;
        }
        __syncthreads();
        const double dg0 = bc[0], tg0 = bc[1];
        const double grad = is_topic ? nu / a + Md * (dg0 - dg) + Es : 0.0;
        const double hinv = is_topic ? -1.0 / (Md * tg + nu / (a * a)) : 0.0;
        double gh = grad * hinv, hs = hinv, gn = grad * grad;
        block_sum3(gh, hs, gn, red);
        const double z = gh / (1.0 / (Md * tg0) + hs);
        const double pd = (grad - z) * hinv;
        double rho = 1.0;
        for (;;) {
            const double mn = block_min(is_topic ? a - rho * pd : 1e300, red);
            if (!(mn < 0.0)) break;
            rho *= 0.5;
        }
        // @finite alpha -= rho * p  (macros.jl:52-54)
        if (is_topic) a = copysign(fmin(fabs(a - rho * pd), 1.7976931348623157e308), a);
        if ((rho * sqrt(gn) < ntol) && (nu / (double)K < ntol)) break;
        nu *= 0.5;
    }
    double a0 = 0.0, sl = 0.0, lin = 0.0;
    if (is_topic) {
        a += TMVB_EPS_D;  // @positive model.alpha
        alpha64[tid] = a;
        alpha32[tid] = fmaxf((float)a, 1.1754944e-38f);
        a0 = a;
        sl = lgamma(a);
        lin = (a - a_estep - TMVB_EPS_D) * Es;
    }
    block_sum3(a0, sl, lin, red);
    return Md * (lgamma(a0) - sl) + lin;
}

// The stand-alone launch: small = [sum_d Elogtheta_d (K_ld) | per-document ELBO terms | sweeps], local = [rowsum (K_ld) | elbo_w].
// want_elbo = 1: the whole ELBO is assembled here, so that an outer iteration needs a single 8-byte read-back; want_elbo = 2:
// only the alpha-dependent part is written (result[1]) -- the kernel then runs beside the M-step kernels on another stream and
// lda_elbo_assemble_kernel adds the M-step's term once both have finished.
__global__ void lda_alpha_kernel(double *__restrict__ alpha64, float *__restrict__ alpha32, const double *__restrict__ small,
                                 const double *__restrict__ local, int K, int K_ld, double Md, int niter, double ntol, int want_elbo,
                                 double *__restrict__ result)
{
    __shared__ double red[3][16];
    __shared__ double bc[2];
    const double Es = threadIdx.x < K ? small[threadIdx.x] : 0.0;
    const double part = lda_alpha_newton(alpha64, alpha32, Es, K, Md, niter, ntol, red, bc);
    if (want_elbo && threadIdx.x == 0) {
        if (want_elbo == 2)
            result[1] = part;
        else
            result[0] = small[K_ld] + part + local[K_ld];
    }
}

// host-callable launch for the other model families that share update_alpha! (fLDA.jl:122-146 is LDA.jl:97-118 verbatim)
int lda_launch_alpha(double *alpha64, float *alpha32, const double *small, int K, int K_ld, double Md, int niter, double ntol, cudaStream_t stream)
{
    const int threads = 32 * ((K + 1 + 31) / 32);
    if (threads > 288) return fail(-2, "update_alpha! on the device supports K <= 287");
    lda_alpha_kernel<<<1, threads, 0, stream>>>(alpha64, alpha32, small, nullptr, K, K_ld, Md, niter, ntol, 0, nullptr);
    TMVB_CUDA(cudaGetLastError());
    return 0;
}

__global__ void lda_elbo_assemble_kernel(const double *__restrict__ small, const double *__restrict__ local, int K_ld, double *__restrict__ result)
{
    result[0] = small[K_ld] + result[1] + local[K_ld];
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

// ------------------------------------------------------------------ fused exchange + M-step (+ update_alpha!) ----------
// One kernel per outer iteration replaces { all-reduce(stats), all-reduce(small), colsum, normalise, update_alpha!, ELBO assembly }
// when the ranks of a box have mapped each other's buffers (tmvb_comm.cuh).  V is cut into `world` row slices; on rank r
//   A. CTA 0 tells every peer that this rank's E-step has finished (flag epoch+1 in the peer's control block); every CTA waits
//      for all peers' flags -- the statistics of all ranks are complete;
//   B. reduce-scatter: each CTA sums its rows of the slice over all ranks with loads from the mapped peer buffers (fixed rank
//      order), accumulates the slice's column sums in fp64 and publishes them (atomics into this rank's partial vector); when
//      the last CTA has arrived CTA 0 raises flag epoch+2 on every rank (itself included);
//   C. every CTA waits for all flags epoch+2, sums the column-sum partials of all ranks (fixed order: bit-identical on every
//      rank), normalises its rows (LDA.jl:121-125), all-gathers them by storing into EVERY rank's next beta buffer, accumulates
//      sum S (ln beta_new - ln beta_old) (Elogpw + the linear part of -Elogqz) and zeroes its local statistics; when the last CTA
//      has arrived CTA 0 raises flag epoch+3 on the peers and waits for theirs -- every slice has landed everywhere;
//   D. CTA 0 sums the ELBO partials of all ranks and assembles the ELBO.
// With do_alpha the grid has one more CTA that does not take part in B and C: after A it sums the small fp64 vectors
// (sum_d Elogtheta_d, per-document ELBO terms, sweeps) of all ranks and runs update_alpha! (LDA.jl:97-118) BESIDE the exchange;
// CTA 0 waits for it before it lets the peers go (their next E-step clears the vector this CTA reads) and before it assembles
// the ELBO.  No grid-wide barrier: two arrival counters and the flag words; no cooperative launch (the grid is at most one CTA
// per SM plus one, resident at once on an otherwise idle device), so the kernel is an ordinary node of the iteration's CUDA graph.
// The epoch and the buffer parity live in the control block: the kernel's parameters do not change from call to call.
// Traffic per rank: (world-1)/world of the table in, the same out, over NVLink; no NCCL call, no host round trip.
struct LdaXchg {
    int V, K, K_ld, rank, world, want_elbo, n_small, do_alpha;
    long long timeout_ns;
    const float *stats[kMaxPeers];
    float *beta_new[kMaxPeers];
    const double *small[kMaxPeers];
    void *ctl[kMaxPeers];
    float *my_stats;
    const float *beta_old;
    double *small_red, *local, *result;
    double *alpha64;
    float *alpha32;
    double Md, ntol;
    int niter;
};

// thread r < world waits until rank r has signalled `target` (self included when with_self); every thread of the CTA returns after it
__device__ __forceinline__ void wait_flags(const CtlView &me, int rank, int world, unsigned long long target, bool with_self, long long timeout_ns)
{
    const int r = threadIdx.x;
    if (r < world && (with_self || r != rank)) {
        const long long t0 = global_ns();
        while (ld_acquire_sys(me.flag + r) < target) {
            if (global_ns() - t0 > timeout_ns) {
                atomicExch(me.status, 1u);
                break;
            }
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void wait_count(unsigned *count, unsigned target, unsigned *status, long long timeout_ns)
{
    const long long t0 = global_ns();
    while (ld_acquire_gpu(count) < target) {
        if (global_ns() - t0 > timeout_ns) {
            atomicExch(status, 2u);
            break;
        }
    }
}

__global__ void __launch_bounds__(256) lda_exchange_mstep_kernel(const LdaXchg x)
{
    extern __shared__ double sh[];  // [RPP][K_ld] column-sum staging | rs[K_ld]
    __shared__ unsigned long long s_epoch[2];
    __shared__ double red[3][16];
    __shared__ double bc[2];
    const int tid = threadIdx.x, G = gridDim.x - x.do_alpha;
    const int K_ld = x.K_ld, CH = K_ld >> 2, RPP = blockDim.x / CH;
    const int rl = tid / CH, c = tid - rl * CH;
    const bool active = rl < RPP;
    double *rs = sh + (size_t)RPP * K_ld;
    void *my_ctl = x.ctl[x.rank];
    const CtlView me = ctl_view(my_ctl);
    if (tid == 0) {
        s_epoch[0] = *me.epoch;
        s_epoch[1] = *me.calls;
    }
    __syncthreads();
    const unsigned long long epoch = s_epoch[0];
    const int parity = (int)(s_epoch[1] & 1);
    double *my_part = me.part + parity * kCtlPartLen;
    const int r0 = (int)((long long)x.V * x.rank / x.world), r1 = (int)((long long)x.V * (x.rank + 1) / x.world);

    if (x.do_alpha && (int)blockIdx.x == G) {
        // ---- the update_alpha! CTA: observer of phase A, then the reduction of `small` and the Newton iteration
        wait_flags(me, x.rank, x.world, epoch + 1, false, x.timeout_ns);
        for (int i = tid; i < x.n_small; i += blockDim.x) {
            double v[kMaxPeers];
#pragma unroll
            for (int pr = 0; pr < kMaxPeers; pr++) v[pr] = pr < x.world ? __ldcg(x.small[pr] + i) : 0.0;
            double a = 0.0;
#pragma unroll
            for (int pr = 0; pr < kMaxPeers; pr++) a += v[pr];
            x.small_red[i] = a;
        }
        __syncthreads();
        const double Es = tid < x.K ? x.small_red[tid] : 0.0;
        __threadfence();
        if (tid == 0) st_release_sys(me.alpha_done, epoch + 1);   // the peers' small vectors have been read
        const double part = lda_alpha_newton(x.alpha64, x.alpha32, Es, x.K, x.Md, x.niter, x.ntol, red, bc);
        if (tid == 0) {
            x.result[1] = part;
            __threadfence();
            st_release_sys(me.alpha_done, epoch + 3);
        }
        return;
    }

    // ---- A: every rank's E-step has finished
    if (blockIdx.x == 0 && tid < x.world && tid != x.rank) {
        __threadfence_system();
        st_release_sys(ctl_view(x.ctl[tid]).flag + x.rank, epoch + 1);
    }
    wait_flags(me, x.rank, x.world, epoch + 1, false, x.timeout_ns);

    // ---- B: reduce-scatter + column sums of the slice
    double cs0 = 0.0, cs1 = 0.0, cs2 = 0.0, cs3 = 0.0;
    if (active) {
        for (int r = r0 + blockIdx.x * RPP + rl; r < r1; r += G * RPP) {
            const size_t q = (size_t)r * K_ld + 4 * c;
            // all peer loads in flight at once (a remote load costs ~2 us: eight dependent ones would be most of the kernel), then
            // the sum in fixed rank order
            float4 v[kMaxPeers];
#pragma unroll
            for (int pr = 0; pr < kMaxPeers; pr++)
                v[pr] = pr < x.world ? __ldcg(reinterpret_cast<const float4 *>(x.stats[pr] + q)) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int pr = 0; pr < kMaxPeers; pr++) {
                acc.x += v[pr].x;
                acc.y += v[pr].y;
                acc.z += v[pr].z;
                acc.w += v[pr].w;
            }
            *reinterpret_cast<float4 *>(x.my_stats + q) = acc;
            cs0 += (double)acc.x;
            cs1 += (double)acc.y;
            cs2 += (double)acc.z;
            cs3 += (double)acc.w;
        }
        double *row = sh + (size_t)rl * K_ld + 4 * c;
        row[0] = cs0;
        row[1] = cs1;
        row[2] = cs2;
        row[3] = cs3;
    }
    __syncthreads();
    if (tid < K_ld) {
        double a = 0.0;
        for (int q = 0; q < RPP; q++) a += sh[(size_t)q * K_ld + tid];
        if (a != 0.0) atomicAdd(my_part + tid, a);
    }
    if (blockIdx.x == 0 && !x.do_alpha)
        for (int i = tid; i < x.n_small; i += blockDim.x) {
            double a = 0.0;
            for (int pr = 0; pr < x.world; pr++) a += __ldcg(x.small[pr] + i);
            x.small_red[i] = a;
        }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        atomicAdd(me.cnt_b, 1u);
    }
    if (blockIdx.x == 0) {
        if (tid == 0) wait_count(me.cnt_b, (unsigned)G, me.status, x.timeout_ns);
        __syncthreads();
        if (tid < x.world) {
            __threadfence_system();
            st_release_sys(ctl_view(x.ctl[tid]).flag + x.rank, epoch + 2);
        }
    }
    wait_flags(me, x.rank, x.world, epoch + 2, true, x.timeout_ns);

    // ---- C: total column sums (fixed rank order), normalise + all-gather the rows, zero the local statistics
    if (tid < K_ld) {
        double v[kMaxPeers];
#pragma unroll
        for (int pr = 0; pr < kMaxPeers; pr++) v[pr] = pr < x.world ? __ldcg(ctl_view(x.ctl[pr]).part + parity * kCtlPartLen + tid) : 0.0;
        double a = 0.0;
#pragma unroll
        for (int pr = 0; pr < kMaxPeers; pr++) a += v[pr];
        rs[tid] = a;
    }
    __syncthreads();
    double eacc = 0.0;
    if (active) {
        double inv[4];
#pragma unroll
        for (int j = 0; j < 4; j++) inv[j] = (4 * c + j < x.K && rs[4 * c + j] > 0.0) ? 1.0 / rs[4 * c + j] : 0.0;
        for (int r = r0 + blockIdx.x * RPP + rl; r < r1; r += G * RPP) {
            const size_t q = (size_t)r * K_ld + 4 * c;
            const float4 S = *reinterpret_cast<const float4 *>(x.my_stats + q);
            float4 b;
            b.x = (float)((double)S.x * inv[0]);
            b.y = (float)((double)S.y * inv[1]);
            b.z = (float)((double)S.z * inv[2]);
            b.w = (float)((double)S.w * inv[3]);
            if (x.want_elbo) {
                const float4 bo = *reinterpret_cast<const float4 *>(x.beta_old + q);
                if (4 * c + 0 < x.K) eacc += (double)(S.x * (logf(b.x + TMVB_EPS) - logf(bo.x + TMVB_EPS)));
                if (4 * c + 1 < x.K) eacc += (double)(S.y * (logf(b.y + TMVB_EPS) - logf(bo.y + TMVB_EPS)));
                if (4 * c + 2 < x.K) eacc += (double)(S.z * (logf(b.z + TMVB_EPS) - logf(bo.z + TMVB_EPS)));
                if (4 * c + 3 < x.K) eacc += (double)(S.w * (logf(b.w + TMVB_EPS) - logf(bo.w + TMVB_EPS)));
            }
            for (int pr = 0; pr < x.world; pr++) *reinterpret_cast<float4 *>(x.beta_new[pr] + q) = b;
            *reinterpret_cast<float4 *>(x.my_stats + q) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    {   // the rows outside the slice (the peers have finished reading them: flag epoch+2)
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 *st4 = reinterpret_cast<float4 *>(x.my_stats);
        const size_t n4 = (size_t)x.V * CH, s0 = (size_t)r0 * CH, s1 = (size_t)r1 * CH;
        for (size_t q = (size_t)blockIdx.x * blockDim.x + tid; q < n4; q += (size_t)G * blockDim.x)
            if (q < s0 || q >= s1) st4[q] = z4;
    }
    if (x.want_elbo) {
        eacc = warp_sum_d(eacc);
        if ((tid & 31) == 0 && eacc != 0.0) atomicAdd(my_part + K_ld, eacc);
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();   // this CTA's stores into the peers' tables before the arrival
        atomicAdd(me.cnt_c, 1u);
    }
    if (blockIdx.x != 0) return;

    // ---- D (CTA 0): every local CTA has stored its rows; the update_alpha! CTA has read the peers' vectors; then the peers
    if (tid == 0) {
        wait_count(me.cnt_c, (unsigned)G, me.status, x.timeout_ns);
        if (x.do_alpha) {
            const long long t0 = global_ns();
            while (ld_acquire_sys(me.alpha_done) < epoch + 1)
                if (global_ns() - t0 > x.timeout_ns) {
                    atomicExch(me.status, 2u);
                    break;
                }
        }
    }
    __syncthreads();
    if (tid < x.world && tid != x.rank) {
        __threadfence_system();
        st_release_sys(ctl_view(x.ctl[tid]).flag + x.rank, epoch + 3);
    }
    wait_flags(me, x.rank, x.world, epoch + 3, false, x.timeout_ns);

    // the reduced ELBO term; rowsum for inspection; recycle the other parity of the partial sums; next call's epoch
    if (tid < K_ld) x.local[tid] = rs[tid];
    if (tid == 0) {
        double a = 0.0;
        for (int pr = 0; pr < x.world; pr++) a += __ldcg(ctl_view(x.ctl[pr]).part + parity * kCtlPartLen + K_ld);
        x.local[K_ld] = a;
        if (x.do_alpha) {
            const long long t0 = global_ns();
            while (ld_acquire_sys(me.alpha_done) < epoch + 3)
                if (global_ns() - t0 > x.timeout_ns) {
                    atomicExch(me.status, 2u);
                    break;
                }
            if (x.want_elbo) x.result[0] = x.small_red[K_ld] + x.result[1] + a;
        }
        *me.cnt_b = 0u;
        *me.cnt_c = 0u;
        *me.epoch = epoch + 3;
        *me.calls = s_epoch[1] + 1;
    }
    double *other = me.part + (parity ^ 1) * kCtlPartLen;
    for (int i = tid; i < kCtlPartLen; i += blockDim.x) other[i] = 0.0;
}

typedef void (*LdaEstepFn)(const LdaDev, int, int, int, int, int *);
// The hybrid kernel is instantiated per K_ld (compile-time strides), one translation unit per value: tmvb_lda_hyb_<K_ld>.cu
// built from tmvb_lda_hyb_inst.cuh; tmvb_lda_hyb.cuh declares the tables.  Other K use the tile kernel.
static const LdaHybLayout *const kLdaHyb[] = {TMVB_LDA_HYB_TABLES};
static const LdaHybLayout *lda_hyb_layout(int lpt, int cpl, int K_ld)
{
    for (const LdaHybLayout *l : kLdaHyb)
        if (l->lpt == lpt && l->cpl == cpl && l->K_ld == K_ld) return l;
    return nullptr;
}
struct LdaPick {
    const void *tile[3];
    const LdaHybLayout *hyb;
    bool elbo;
};
static const void *lda_pick(const Bucket &b, const void *ctx)
{
    const LdaPick *pk = static_cast<const LdaPick *>(ctx);
    if (b.hyb) {
        for (int v = 0; v < kNumHybVariants; v++)
            if (kHybVariant[v][0] == b.warps && kHybVariant[v][1] == b.nr && kHybVariant[v][2] == (b.cap > 0 ? 1 : 0))
                return (const void *)pk->hyb->fn[pk->elbo ? 1 : 0][v];
        return nullptr;
    }
    return pk->tile[b.warps >= 4 ? 2 : b.warps - 1];
}
#define TMVB_LDA_FN1(L, C) {(LdaEstepFn)lda_estep_kernel<L, C, 1, false>, (LdaEstepFn)lda_estep_kernel<L, C, 1, true>},
#define TMVB_LDA_FN2(L, C) {(LdaEstepFn)lda_estep_kernel<L, C, 2, false>, (LdaEstepFn)lda_estep_kernel<L, C, 2, true>},
#define TMVB_LDA_FN4(L, C) {(LdaEstepFn)lda_estep_kernel<L, C, 4, false>, (LdaEstepFn)lda_estep_kernel<L, C, 4, true>},
// [warps per document - 1][lane layout][want_elbo]
// [warps per document: 1, 2, 4][lane layout][want_elbo]  (eight warps per document measured slower than four at K=200: 25 vs 17 ms)
static const LdaEstepFn kLdaEstep[3][kNumLaneLayouts][2] = {{TMVB_FOR_EACH_LAYOUT(TMVB_LDA_FN1)}, {TMVB_FOR_EACH_LAYOUT(TMVB_LDA_FN2)},
                                                            {TMVB_FOR_EACH_LAYOUT(TMVB_LDA_FN4)}};

}  // namespace tmvb

using namespace tmvb;

struct tmvb_lda_s {
    Shard s;
    bool params_set = false, elbo_valid = false;
    bool beta_fresh = false, E_fresh = false;   // beta_old == beta / Elogtheta_old == Elogtheta on the device (set by upload, cleared by any step)
    float *d_alpha = nullptr;
    float *d_Elogtheta = nullptr, *d_Elogtheta_old = nullptr, *d_gamma = nullptr;
    std::vector<double> h_alpha;        // fp64 master copy of alpha (update_alpha! runs in fp64 on the host)
    double *d_alpha64 = nullptr;        // [K_ld] fp64 alpha on the device (master while alpha_on_device)
    bool alpha_on_device = false, elbo_dev_valid = false;
    double *d_small = nullptr;          // [K_ld+2], summed over ranks
    double *d_local = nullptr;          // [K_ld] rowsum | [K_ld] elbo_w | [K_ld+1] scratch for the standalone ELBO
    bool no_scatter = false;            // predict: the E-step leaves the statistics alone
    bool solo = true;                   // no other rank: nothing sums `small` between the E-step and update_alpha!
    std::vector<Shard::LaunchGraph> iter_graphs;   // captured outer iterations (tmvb_lda_iterate), keyed like the E-step graphs
    Comm comm;                          // peer-memory exchange (multi-GPU), see tmvb_comm.cuh
    // host mirror (tmvb_lda_arm_host_mirror): the caller's page-locked Elogtheta / gamma arrays as the device sees them, armed for
    // the next E-step; mirror_valid = those arrays hold the rows of the last E-step (tmvb_lda_download then skips their transfer)
    float *mirror_E = nullptr, *mirror_gamma = nullptr, *mirror_E_host = nullptr, *mirror_gamma_host = nullptr;
    bool mirror_armed = false, mirror_valid = false;
};

namespace {

// the armed host mirror rides with the next E-step only (the rows every other E-step writes would be overwritten anyway)
void take_mirror(tmvb_lda_t h, LdaDev *p)
{
    h->mirror_valid = false;
    if (!h->mirror_armed) return;
    h->mirror_armed = false;
    if (h->no_scatter) return;
    p->host_E = h->mirror_E;
    p->host_gamma = h->mirror_gamma;
    p->perm = h->s.d_perm;
    h->mirror_valid = true;
    h->s.st.d2h_bytes += 2 * h->s.M * h->s.K * 4;   // written over the bus by the E-step kernels
}

LdaDev dev_view(tmvb_lda_t h)
{
    Shard &s = h->s;
    LdaDev p;
    memset(&p, 0, sizeof(p));  // the struct is also the key of the captured launch graph: no indeterminate padding
    p.K = (int)s.K;
    p.K_ld = s.K_ld;
    p.V = (int)s.V;
    p.RS = s.RS;
    p.M = s.M;
    p.beta = s.d_beta[s.cur];
    p.alpha = h->d_alpha;
    p.stats = s.d_stats;
    p.doc_off = s.d_doc_off;
    p.terms = s.d_terms;
    p.counts = s.d_counts;
    p.doc_c = s.d_doc_c;
    p.Elogtheta = h->d_Elogtheta;
    p.Elogtheta_old = h->d_Elogtheta_old;
    p.gamma = h->d_gamma;
    p.small = h->d_small;
    p.viter = 0;
    p.vtol = 0.f;
    p.stage_bulk = env_int("TMVB_STAGE_BULK", 1);
    p.dbg = env_int("TMVB_DBG", 0) | (h->no_scatter ? 1 : 0);   // bit 0: the scatter pass computes but does not store (predict)
    return p;
}

void lda_free(tmvb_lda_t h)
{
    cudaSetDevice(h->s.device);
    if (h->s.stream) cudaStreamSynchronize(h->s.stream);
    for (auto &g : h->iter_graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    h->iter_graphs.clear();
    cudaFree(h->d_alpha);
    cudaFree(h->d_alpha64);
    cudaFree(h->d_Elogtheta);
    cudaFree(h->d_Elogtheta_old);
    cudaFree(h->d_gamma);
    cudaFree(h->d_small);
    cudaFree(h->d_local);
    comm_free(&h->comm);
    shard_free(&h->s);
}

// bring the fp64 host copy of alpha up to date after a device-side update_alpha!
int sync_alpha(tmvb_lda_t h)
{
    if (!h->alpha_on_device) return 0;
    Shard &s = h->s;
    TMVB_CUDA(cudaMemcpyAsync(s.h_pinned, h->d_alpha64, s.K * 8, cudaMemcpyDeviceToHost, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    memcpy(h->h_alpha.data(), s.h_pinned, s.K * 8);
    s.st.d2h_bytes += s.K * 8;
    h->alpha_on_device = false;
    return 0;
}

// a non-zero time-out status of the peer exchange: report it, and clear the flag so that the next exchange starts clean
int comm_raise(tmvb_lda_t h, unsigned st)
{
    if (!st) return 0;
    cudaMemsetAsync(h->comm.d_ctl + 132, 0, 4, h->s.stream);
    return fail(900 + (int)st, "peer exchange timed out (%s): a rank did not reach tmvb_lda_exchange_mstep within %d ms; the statistics of this iteration are incomplete",
                st == 1 ? "peer barrier" : "grid barrier", h->comm.timeout_ms);
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
    tmvb_lda_t h = new tmvb_lda_s();
    int rc = shard_create(&h->s, K, M, V, device, stream, 3 * ((K + 7) / 8 * 8) + 8);
    if (rc == 0) {
        Shard &s = h->s;
        const size_t km = (size_t)std::max<int64_t>(M, 1) * s.K_ld;
        cudaError_t e = cudaSuccess;
        auto A = [&](void **p, size_t bytes) {
            if (e == cudaSuccess) e = cudaMalloc(p, bytes);
            if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes, s.stream);
        };
        A((void **)&h->d_alpha, s.K_ld * 4);
        A((void **)&h->d_alpha64, s.K_ld * 8);
        A((void **)&h->d_Elogtheta, km * 4);
        A((void **)&h->d_Elogtheta_old, km * 4);
        A((void **)&h->d_gamma, km * 4);
        A((void **)&h->d_small, (s.K_ld + 2) * 8);
        A((void **)&h->d_local, (3 * s.K_ld + 2) * 8);
        // opt in to the large dynamic shared memory for both instantiations of this K
        for (int w = 0; w < 3; w++)
            for (int eb = 0; eb < 2 && e == cudaSuccess; eb++)
                e = cudaFuncSetAttribute((const void *)kLdaEstep[w][s.layout][eb], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem_optin);
        if (e != cudaSuccess) rc = fail((int)e, "device allocation failed: %s", cudaGetErrorString(e));
    }
    if (rc != 0) {
        lda_free(h);
        delete h;
        return rc;
    }
    h->h_alpha.assign(K, 1.0);
    {  // alpha = ones(K) (gpuLDA.jl:55) on the device as well
        std::vector<float> ones(K, 1.0f);
        rc = tmvb_lda_set_alpha(h, ones.data());
        if (rc != 0) {
            lda_free(h);
            delete h;
            return rc;
        }
    }
    *out = h;
    return 0;
}

int tmvb_lda_destroy(tmvb_lda_t h)
{
    if (!h) return 0;
    lda_free(h);
    delete h;
    return 0;
}

int tmvb_lda_kld(tmvb_lda_t h, int64_t *K_ld)
{
    TMVB_CHECK_ARG(h && K_ld, "NULL argument");
    *K_ld = h->s.K_ld;
    return 0;
}

static int lda_set_corpus(tmvb_lda_t h, const int64_t *N_cumsum, const void *terms, const void *counts, int elem_bytes);

int tmvb_lda_set_corpus(tmvb_lda_t h, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts)
{
    return lda_set_corpus(h, N_cumsum, terms, counts, 8);
}

int tmvb_lda_set_corpus32(tmvb_lda_t h, const int64_t *N_cumsum, const int32_t *terms, const int32_t *counts)
{
    return lda_set_corpus(h, N_cumsum, terms, counts, 4);
}

static int lda_set_corpus(tmvb_lda_t h, const int64_t *N_cumsum, const void *terms, const void *counts, int elem_bytes)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    // Warps per document in the tile kernel.  Small tiles (K_ld <= 64): one -- two measured the same E-step time (2.93 vs
    // 2.89 ms at NSF K=50), the shorter per-document latency is paid for by fewer documents in flight.  Large tiles (K=200:
    // 65 KB for an 80-token document, three documents per SM): the tile bounds the residency, so more warps per
    // document are free parallelism (K=200, 200k documents: 34.8 / 26.2 / 17.7 ms with one / two / four warps; 25 ms with eight).
    const int force_w = env_int("TMVB_LDA_WARPS", 0);
    const int tile_w = force_w ? (force_w >= 4 ? 4 : std::max(1, force_w)) : (s.K_ld > 64 ? 4 : 1);
    TMVB_TRY(shard_set_corpus(&s, N_cumsum, terms, counts, lda_fixed_smem(s.RS, s.lpt, std::max(2, tile_w)), elem_bytes));
    const size_t per_tok = (size_t)s.RS * 4 + 8;
    for (Bucket &b : s.buckets) {
        b.warps = tile_w;
        b.hyb = 0;
        b.smem = lda_fixed_smem(s.RS, s.lpt, b.warps) + (size_t)b.cap * per_tok;
        b.grid = 0;
    }
    // Hybrid kernel (default): classes (W, NR, tile capacity) by ascending document length.  The tile capacity of a class is
    // what leaves room for the number of resident CTAs its register allocation allows (12 / W: 168 registers per thread),
    // then for fewer; documents longer than the last class stay with the tile kernel (which reads overflow rows from L2).
    const LdaHybLayout *hl = lda_hyb_layout(s.lpt, s.cpl, s.K_ld);
    if (hl && env_int("TMVB_LDA_HYB", 1)) {
        const int S = 32 / s.lpt, M = (int)s.len_sorted.size();
        struct Cls {
            int W, NR, cap, maxlen;
        };
        std::vector<Cls> cls;
        auto cap_for = [&](int W, int occ) {
            // 228 KB of shared memory per SM, 1 KB of which every resident CTA reserves for the system
            const long long budget = (long long)(228 * 1024 - 1024 * occ) / occ;
            const long long room = std::min<long long>(budget, (long long)s.smem_optin) - (long long)lda_hyb_fixed_smem(s.K_ld, s.lpt, W);
            return room <= 0 ? 0 : (int)(room / (long long)per_tok) / 4 * 4;   // cnt_s / term_s stay 16-byte aligned
        };
        // class list "W:NR:occ,..." (occ = resident CTAs per SM the tile capacity leaves room for; 0 = no tile), by ascending length
        const char *spec = getenv("TMVB_HYB_CLASSES");
        // (measured on B200, NSF K=50 and cfg4 K=200: the more of a document one warp holds in registers the better -- instructions per
        // document count for more than resident warps -- so one warp takes documents up to 6 rounds, two warps up to 12, four up to 24)
        // a small shard (a quarter of NSF or less per GPU) takes fewer, wider classes: every launch has a tail, and at 32 k documents
        // seven launches beat eleven (E-step 0.346 vs 0.370 ms; equal at 16 k, the eleven win from 64 k up)
        if (!spec || !*spec)
            spec = s.K_ld > 64 ? "4:3:0,4:4:0,4:5:0,4:6:0,4:6:2,4:6:1"
                   : M < 48000 ? "1:2:0,1:4:0,1:6:0,2:4:0,2:6:0,4:6:0,4:6:1"
                               : "1:2:0,1:3:0,1:4:0,1:5:0,1:6:0,2:4:0,2:5:0,2:6:0,4:5:0,4:6:0,4:6:1";
        for (const char *q = spec; *q;) {
            int W = 0, NR = 0, occ = 0;
            if (sscanf(q, "%d:%d:%d", &W, &NR, &occ) != 3) return fail(-1, "invalid argument: TMVB_HYB_CLASSES must be W:NR:occ[,W:NR:occ...]");
            bool known = false;
            for (int v = 0; v < kNumHybVariants; v++)
                known = known || (kHybVariant[v][0] == W && kHybVariant[v][1] == NR && kHybVariant[v][2] == (occ > 0 ? 1 : 0) && hl->fn[0][v]);
            if (!known) return fail(-1, "invalid argument: TMVB_HYB_CLASSES names a variant that is not built for this K (W=%d NR=%d)", W, NR);
            cls.push_back({W, NR, occ > 0 ? cap_for(W, occ) : 0, 0});
            while (*q && *q != ',') q++;
            if (*q == ',') q++;
        }
        for (Cls &c : cls) c.maxlen = c.W * c.NR * S + c.cap;
        const int limit = cls.back().maxlen;
        int first = 0;  // documents are sorted by length, longest first
        while (first < M && s.len_sorted[first] > limit) first++;
        std::vector<Bucket> nb;
        for (Bucket b : s.buckets) {
            if (b.doc_begin >= first) continue;
            b.doc_end = std::min(b.doc_end, first);
            nb.push_back(b);
        }
        int begin = first;
        for (int ci = (int)cls.size() - 1; ci >= 0 && begin < M; ci--) {
            const int lo = ci > 0 ? cls[ci - 1].maxlen : -1;
            int end = begin;
            while (end < M && s.len_sorted[end] > lo) end++;
            if (end == begin) continue;
            Bucket b;
            b.doc_begin = begin;
            b.doc_end = end;
            // the tile only needs to hold the longest document of the launch
            b.cap = cls[ci].cap > 0 ? std::max(4, std::min(cls[ci].cap / 4 * 4, (s.len_sorted[begin] - cls[ci].W * cls[ci].NR * S + 3) / 4 * 4)) : 0;
            b.cap2 = 0;
            b.warps = cls[ci].W;
            b.nr = cls[ci].NR;
            b.hyb = 1;
            b.smem = lda_hyb_fixed_smem(s.K_ld, s.lpt, b.warps) + (size_t)b.cap * per_tok;
            b.grid = 0;
            nb.push_back(b);
            begin = end;
        }
        if ((int)nb.size() > kMaxBuckets) return fail(-1, "internal: too many launch buckets");
        s.buckets.swap(nb);
        for (int eb = 0; eb < 2; eb++)
            for (int v = 0; v < kNumHybVariants; v++)
                if (hl->fn[eb][v])
                    TMVB_CUDA(cudaFuncSetAttribute((const void *)hl->fn[eb][v], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem_optin));
        return 0;
    }
    return 0;
}

int tmvb_lda_set_alpha(tmvb_lda_t h, const float *alpha)
{
    TMVB_CHECK_ARG(h && alpha, "NULL argument");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    for (int64_t i = 0; i < s.K; i++) {
        TMVB_CHECK_ARG(alpha[i] > 0.f && isfinite(alpha[i]), "alpha must be positive and finite");  // modelutils.jl:262-263
        h->h_alpha[i] = (double)alpha[i];
    }
    std::vector<float> pad(s.K_ld, 0.f);
    memcpy(pad.data(), alpha, s.K * 4);
    TMVB_CUDA(cudaMemcpyAsync(h->d_alpha, pad.data(), s.K_ld * 4, cudaMemcpyHostToDevice, s.stream));
    TMVB_CUDA(cudaMemcpyAsync(h->d_alpha64, h->h_alpha.data(), s.K * 8, cudaMemcpyHostToDevice, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    s.st.h2d_bytes += s.K * 12;
    h->alpha_on_device = false;
    h->elbo_dev_valid = false;
    return 0;
}

int tmvb_lda_upload(tmvb_lda_t h, const float *alpha, const float *beta, const float *Elogtheta, const float *gamma)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    if (alpha) TMVB_TRY(tmvb_lda_set_alpha(h, alpha));
    if (beta && s.V > 0) {
        TMVB_TRY(shard_upload_rows(&s, beta, s.d_beta[s.cur], s.V, nullptr, 0));
        TMVB_TRY(shard_check_stochastic(&s, s.d_beta[s.cur]));   // the row sums of check_model, on the device copy
        // beta_old = copy(beta)  (LDA.jl:36)
        TMVB_CUDA(cudaMemcpyAsync(s.d_beta[s.cur ^ 1], s.d_beta[s.cur], (size_t)s.V * s.K_ld * 4, cudaMemcpyDeviceToDevice, s.stream));
        h->beta_fresh = true;
    }
    if ((Elogtheta || gamma) && s.M > 0) {
        TMVB_CHECK_ARG(s.corpus_set, "set_corpus must precede the upload of per-document parameters");
        h->mirror_valid = false;
        TMVB_TRY(shard_upload_rows(&s, Elogtheta, h->d_Elogtheta, s.M, s.d_perm, 1));
        // Elogtheta_old = deepcopy(Elogtheta)  (LDA.jl:39)
        if (Elogtheta)
            TMVB_CUDA(cudaMemcpyAsync(h->d_Elogtheta_old, h->d_Elogtheta, (size_t)s.M * s.K_ld * 4, cudaMemcpyDeviceToDevice, s.stream));
        if (Elogtheta) h->E_fresh = true;
        TMVB_TRY(shard_upload_rows(&s, gamma, h->d_gamma, s.M, s.d_perm, 2));
    }
    int verr = 0;
    TMVB_TRY(shard_validation(&s, &verr));
    // the messages of check_model(::gpuLDA), modelutils.jl:264-273
    if (verr & 0x4003) return fail(-5, "beta must be a right stochastic matrix.");
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
    Shard &s = h->s;
    TMVB_CHECK_ARG(s.corpus_set, "set_corpus has not been called");
    TMVB_CUDA(cudaSetDevice(s.device));
    LdaDev p = dev_view(h);
    p.viter = viter;
    p.vtol = vtol;
    take_mirror(h, &p);
    TMVB_CUDA(cudaEventRecord(s.ev[0], s.stream));
    TMVB_CUDA(cudaMemsetAsync(h->d_small, 0, (s.K_ld + 2) * 8, s.stream));
    h->elbo_dev_valid = false;
    LdaPick pk;
    pk.tile[0] = (const void *)kLdaEstep[0][s.layout][want_elbo != 0];
    pk.tile[1] = (const void *)kLdaEstep[1][s.layout][want_elbo != 0];
    pk.tile[2] = (const void *)kLdaEstep[2][s.layout][want_elbo != 0];
    pk.hyb = lda_hyb_layout(s.lpt, s.cpl, s.K_ld);
    pk.elbo = want_elbo != 0;
    TMVB_TRY(shard_launch(&s, lda_pick, &pk, &p, sizeof(p)));
    TMVB_CUDA(cudaEventRecord(s.ev[1], s.stream));
    s.estep_timed = true;
    h->elbo_valid = (want_elbo != 0);
    h->E_fresh = false;
    return 0;
}

// the inner loop of predict (modelutils.jl:846-855): the E-step without update_beta!(model, d) -- no statistics are scattered
int tmvb_lda_predict(tmvb_lda_t h, int viter, float vtol)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    h->no_scatter = true;
    const int rc = tmvb_lda_estep(h, viter, vtol, 0);
    h->no_scatter = false;
    return rc;
}

int tmvb_lda_reduce_buffers(tmvb_lda_t h, void **stats, int64_t *n_stats, void **small, int64_t *n_small)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    h->solo = false;   // the caller sums these buffers over ranks between estep and mstep
    if (stats) *stats = h->s.d_stats;
    if (n_stats) *n_stats = (int64_t)h->s.V * h->s.K_ld;
    if (small) *small = h->d_small;
    if (n_small) *n_small = h->s.K_ld + 2;
    return 0;
}

int tmvb_lda_mstep(tmvb_lda_t h)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    TMVB_CUDA(cudaEventRecord(s.ev[2], s.stream));
    TMVB_TRY(shard_normalize(&s, h->d_local, h->elbo_valid, true));
    TMVB_CUDA(cudaEventRecord(s.ev[3], s.stream));
    s.mstep_timed = true;
    h->beta_fresh = false;
    return 0;
}

int tmvb_lda_comm_export(tmvb_lda_t h, void *blob, int64_t blob_bytes)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CHECK_ARG(blob_bytes >= TMVB_COMM_BLOB_BYTES, "blob must hold TMVB_COMM_BLOB_BYTES");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    void *bufs[kCommBufs] = {s.d_stats, s.d_beta[0], s.d_beta[1], h->d_small, nullptr};
    return comm_export(&h->comm, bufs, blob, (size_t)blob_bytes);
}

int tmvb_lda_comm_connect(tmvb_lda_t h, int rank, int world, const void *blobs, int64_t blob_bytes)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CUDA(cudaSetDevice(h->s.device));
    h->solo = false;
    return comm_connect(&h->comm, rank, world, blobs, (size_t)blob_bytes);
}

// enqueue the fused exchange kernel on s.stream; do_alpha: the kernel also runs update_alpha! and assembles the ELBO
static int lda_enqueue_exchange(tmvb_lda_t h, bool do_alpha, int64_t M_total, int niter, double ntol)
{
    Shard &s = h->s;
    Comm &c = h->comm;
    LdaXchg x;
    memset(&x, 0, sizeof(x));
    x.V = (int)s.V;
    x.K = (int)s.K;
    x.K_ld = s.K_ld;
    x.rank = c.rank;
    x.world = c.world;
    x.want_elbo = h->elbo_valid ? 1 : 0;
    x.n_small = s.K_ld + 2;
    x.do_alpha = do_alpha ? 1 : 0;
    x.timeout_ns = (long long)c.timeout_ms * 1000000ll;
    const int nb = s.cur ^ 1;  // the buffer that becomes `beta`
    for (int r = 0; r < kMaxPeers; r++) {
        const bool ok = r < c.world;
        x.stats[r] = ok ? (const float *)c.peer[0][r] : nullptr;
        x.beta_new[r] = ok ? (float *)c.peer[1 + nb][r] : nullptr;
        x.small[r] = ok ? (const double *)c.peer[3][r] : nullptr;
        x.ctl[r] = ok ? c.peer[4][r] : nullptr;
    }
    x.my_stats = s.d_stats;
    x.beta_old = s.d_beta[s.cur];
    x.small_red = c.d_small_red;
    x.local = h->d_local;
    x.result = h->d_local + 2 * s.K_ld + 1;
    x.alpha64 = h->d_alpha64;
    x.alpha32 = h->d_alpha;
    x.Md = (double)M_total;
    x.niter = niter;
    x.ntol = ntol;
    const int CH = s.K_ld / 4, RPP = 256 / CH;
    const size_t smem = ((size_t)RPP * s.K_ld + s.K_ld) * 8;
    const int rows = (int)((s.V + c.world - 1) / c.world);
    const int G = std::min(s.n_sm, std::max(1, (rows + RPP - 1) / RPP));
    // every CTA of the grid spins on its siblings: they must all be resident at once
    int occ = 0;
    TMVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)lda_exchange_mstep_kernel, 256, smem));
    if ((long long)occ * s.n_sm < G + x.do_alpha) return fail(-4, "internal: the exchange kernel's grid cannot be co-resident");
    lda_exchange_mstep_kernel<<<G + x.do_alpha, 256, smem, s.stream>>>(x);
    TMVB_CUDA(cudaGetLastError());
    // the reduced small vector replaces the local one: no peer reads the local one any more once this kernel has finished
    // (a peer lets this rank go only after its update_alpha! CTA / its CTA 0 has read it)
    TMVB_CUDA(cudaMemcpyAsync(h->d_small, c.d_small_red, (s.K_ld + 2) * 8, cudaMemcpyDeviceToDevice, s.stream));
    return 0;
}

int tmvb_lda_exchange_mstep(tmvb_lda_t h)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    Comm &c = h->comm;
    TMVB_CHECK_ARG(c.connected, "tmvb_lda_comm_connect has not been called");
    TMVB_CHECK_ARG(s.K_ld + 1 <= kCtlPartLen, "K too large for the exchange control block");
    TMVB_CUDA(cudaSetDevice(s.device));
    TMVB_CUDA(cudaEventRecord(s.ev[2], s.stream));
    h->beta_fresh = false;
    if (s.V > 0) {
        TMVB_TRY(lda_enqueue_exchange(h, false, 0, 0, 0.0));
        s.st.kernel_launches++;
        s.cur ^= 1;  // beta_old <- beta ; beta <- new  (LDA.jl:122-123)
    }
    TMVB_CUDA(cudaEventRecord(s.ev[3], s.stream));
    s.mstep_timed = true;
    return 0;
}

// One outer iteration of train! (gpuLDA.jl:355-371: the folded inner loop, update_beta!, update_alpha!, and the ELBO that
// check_elbo! reads) as ONE CUDA graph launch: E-step bucket launches on four streams -> [one GPU: colsum + normalise with
// update_alpha! beside them | several GPUs: the fused exchange kernel, which also runs update_alpha!] -> ELBO read-back into
// page-locked memory.  The graph is captured once per (beta buffer parity, keyword set); an iteration then costs the host one
// cudaGraphLaunch and -- only when the ELBO is wanted -- one stream synchronisation.
int tmvb_lda_iterate(tmvb_lda_t h, int viter, float vtol, int want_elbo, int64_t M_total, int niter, double ntol, double *elbo)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CHECK_ARG(viter >= 1, "viter must be at least 1");
    TMVB_CHECK_ARG(vtol >= 0.f && ntol >= 0.0 && niter >= 0, "iteration/tolerance parameters must be nonnegative");  // gpuLDA.jl:349-350
    TMVB_CHECK_ARG(!want_elbo || elbo != nullptr, "elbo pointer is NULL");
    Shard &s = h->s;
    Comm &c = h->comm;
    TMVB_CHECK_ARG(s.corpus_set, "set_corpus has not been called");
    TMVB_CHECK_ARG(h->solo || c.connected, "tmvb_lda_iterate needs one GPU or connected peers (the NCCL path sums the buffers between estep and mstep)");
    TMVB_CHECK_ARG(!c.connected || s.K_ld + 1 <= kCtlPartLen, "K too large for the exchange control block");
    TMVB_CUDA(cudaSetDevice(s.device));
    const bool fuse_alpha = c.connected && s.K + 1 <= 256 && s.V > 0;
    LdaDev p = dev_view(h);
    p.viter = viter;
    p.vtol = vtol;
    take_mirror(h, &p);
    LdaPick pk;
    pk.tile[0] = (const void *)kLdaEstep[0][s.layout][want_elbo != 0];
    pk.tile[1] = (const void *)kLdaEstep[1][s.layout][want_elbo != 0];
    pk.tile[2] = (const void *)kLdaEstep[2][s.layout][want_elbo != 0];
    pk.hyb = lda_hyb_layout(s.lpt, s.cpl, s.K_ld);
    pk.elbo = want_elbo != 0;
    std::string key;
    TMVB_TRY(shard_launch_key(&s, lda_pick, &pk, &p, sizeof(p), &key));
    {
        const long long extra[6] = {want_elbo != 0, (long long)M_total, niter, c.connected ? c.world : 0, c.rank, 0};
        key.append((const char *)extra, sizeof(extra));
        key.append((const char *)&ntol, sizeof(ntol));
        key.append("iter", 4);
    }
    h->elbo_valid = want_elbo != 0;   // read by the enqueue helpers below
    h->beta_fresh = h->E_fresh = false;
    double *result = h->d_local + 2 * s.K_ld + 1;
    const int threads = 32 * (((int)s.K + 1 + 31) / 32);

    // timing events inside a capture must be recorded as external event nodes to be readable afterwards
    auto rec = [&](cudaEvent_t ev) -> cudaError_t {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(s.stream, &cs);
        return cs == cudaStreamCaptureStatusActive ? cudaEventRecordWithFlags(ev, s.stream, cudaEventRecordExternal) : cudaEventRecord(ev, s.stream);
    };
    const int64_t launches0 = s.st.kernel_launches;
    auto enqueue = [&]() -> int {
        TMVB_CUDA(rec(s.ev[0]));
        TMVB_CUDA(cudaMemsetAsync(h->d_small, 0, (s.K_ld + 2) * 8, s.stream));
        if (!s.buckets.empty()) TMVB_TRY(shard_enqueue_buckets(&s, lda_pick, &pk, &p));
        TMVB_CUDA(rec(s.ev[1]));
        TMVB_CUDA(rec(s.ev[2]));
        if (c.connected && s.V > 0) {
            TMVB_TRY(lda_enqueue_exchange(h, fuse_alpha, M_total, niter, ntol));
            if (!fuse_alpha) {
                lda_alpha_kernel<<<1, threads, 0, s.stream>>>(h->d_alpha64, h->d_alpha, h->d_small, h->d_local, (int)s.K, s.K_ld, (double)M_total, niter, ntol,
                                                              want_elbo ? 1 : 0, result);
                TMVB_CUDA(cudaGetLastError());
            }
        } else {
            // update_alpha! on an auxiliary stream beside colsum / normalise (it needs only sum_d Elogtheta_d)
            const bool beside = s.n_streams > 1;
            cudaStream_t as = beside ? s.aux[0] : s.stream;
            if (beside) {
                TMVB_CUDA(cudaEventRecord(s.ev_fork, s.stream));
                TMVB_CUDA(cudaStreamWaitEvent(as, s.ev_fork, 0));
            }
            lda_alpha_kernel<<<1, threads, 0, as>>>(h->d_alpha64, h->d_alpha, h->d_small, h->d_local, (int)s.K, s.K_ld, (double)M_total, niter, ntol,
                                                    want_elbo ? 2 : 0, result);
            TMVB_CUDA(cudaGetLastError());
            if (s.V > 0) {
                TMVB_TRY(shard_normalize(&s, h->d_local, want_elbo != 0, true));
                s.cur ^= 1;   // shard_normalize flipped the buffers: the flip is redone once per LAUNCH below, not per capture
            }
            if (beside) {
                TMVB_CUDA(cudaEventRecord(s.ev_join[0], as));
                TMVB_CUDA(cudaStreamWaitEvent(s.stream, s.ev_join[0], 0));
            }
            if (want_elbo) {
                lda_elbo_assemble_kernel<<<1, 1, 0, s.stream>>>(h->d_small, h->d_local, s.K_ld, result);
                TMVB_CUDA(cudaGetLastError());
            }
        }
        TMVB_CUDA(rec(s.ev[3]));
        if (want_elbo) {
            TMVB_CUDA(cudaMemcpyAsync(s.h_pinned, result, 8, cudaMemcpyDeviceToHost, s.stream));
            if (c.connected) TMVB_CUDA(cudaMemcpyAsync(s.h_pinned + 1, c.d_ctl + 128, 8, cudaMemcpyDeviceToHost, s.stream));
        }
        return 0;
    };

    cudaGraphExec_t exec = nullptr;
    for (auto &g : h->iter_graphs)
        if (g.key == key) exec = g.exec;
    if (!exec && s.use_graphs) {
        if (h->iter_graphs.size() >= 8) {
            for (auto &g : h->iter_graphs) cudaGraphExecDestroy(g.exec);
            h->iter_graphs.clear();
        }
        cudaGraph_t graph = nullptr;
        TMVB_CUDA(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue();
        const cudaError_t e = cudaStreamEndCapture(s.stream, &graph);
        if (rc == 0 && e == cudaSuccess && graph) {
            Shard::LaunchGraph lg;
            lg.key = key;
            if (cudaGraphInstantiate(&lg.exec, graph, 0) == cudaSuccess) {
                h->iter_graphs.push_back(lg);
                exec = lg.exec;
            }
        }
        if (graph) cudaGraphDestroy(graph);
        if (!exec) {
            cudaGetLastError();
            s.use_graphs = false;   // capture is not possible here (the caller's stream is capturing, ...): launch directly
            if (rc != 0) return rc;
        }
    }
    if (exec)
        TMVB_CUDA(cudaGraphLaunch(exec, s.stream));
    else
        TMVB_TRY(enqueue());
    // host-side bookkeeping of what the graph did
    s.st.kernel_launches = launches0 + (int64_t)s.buckets.size() + (c.connected ? (fuse_alpha ? 1 : 2) : (s.V > 0 ? 3 : 1) + (want_elbo ? 1 : 0));
    if (s.V > 0) s.cur ^= 1;  // beta_old <- beta ; beta <- new  (LDA.jl:122-123)
    s.estep_timed = s.mstep_timed = true;
    h->alpha_on_device = true;
    h->elbo_dev_valid = h->elbo_valid;
    if (want_elbo) {
        TMVB_CUDA(cudaStreamSynchronize(s.stream));
        s.st.d2h_bytes += c.connected ? 16 : 8;
        if (c.connected) TMVB_TRY(comm_raise(h, reinterpret_cast<const unsigned *>(s.h_pinned + 1)[1]));
        *elbo = s.h_pinned[0];
    }
    return 0;
}

int tmvb_lda_comm_status(tmvb_lda_t h, int *status)
{
    TMVB_CHECK_ARG(h && status, "NULL argument");
    *status = 0;
    if (!h->comm.d_ctl) return 0;
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    unsigned st = 0;
    TMVB_CUDA(cudaMemcpyAsync(&st, h->comm.d_ctl + 132, 4, cudaMemcpyDeviceToHost, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    *status = (int)st;
    return comm_raise(h, st);
}

int tmvb_lda_get_elogtheta_sum(tmvb_lda_t h, double *out)
{
    TMVB_CHECK_ARG(h && out, "NULL argument");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    TMVB_CUDA(cudaMemcpyAsync(s.h_pinned, h->d_small, (s.K_ld + 2) * 8, cudaMemcpyDeviceToHost, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    memcpy(out, s.h_pinned, s.K * 8);
    s.st.d2h_bytes += (s.K_ld + 2) * 8;
    return 0;
}

int tmvb_lda_update_alpha(tmvb_lda_t h, int64_t M_total, int niter, double ntol, float *alpha_out)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CHECK_ARG(niter >= 0 && ntol >= 0.0, "iteration/tolerance parameters must be nonnegative");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    // LDA.jl:97-118 on the device in fp64 (one warp), fused with the ELBO assembly; asynchronous unless alpha_out is given
    double *result = h->d_local + 2 * s.K_ld + 1;   // [0] the ELBO, [1] its alpha-dependent part
    const int threads = 32 * (((int)s.K + 1 + 31) / 32);
    if (h->solo && s.estep_timed && s.n_streams > 1) {
        // one GPU: the Newton iteration needs only sum_d Elogtheta_d (ready when the E-step is), so it runs on an auxiliary
        // stream beside colsum / normalize; the ELBO is assembled on the main stream when both are done
        TMVB_CUDA(cudaStreamWaitEvent(s.aux[0], s.ev[1], 0));
        lda_alpha_kernel<<<1, threads, 0, s.aux[0]>>>(h->d_alpha64, h->d_alpha, h->d_small, h->d_local, (int)s.K, s.K_ld, (double)M_total, niter, ntol,
                                                      h->elbo_valid ? 2 : 0, result);
        TMVB_CUDA(cudaGetLastError());
        TMVB_CUDA(cudaEventRecord(s.ev_join[0], s.aux[0]));
        TMVB_CUDA(cudaStreamWaitEvent(s.stream, s.ev_join[0], 0));
        if (h->elbo_valid) {
            lda_elbo_assemble_kernel<<<1, 1, 0, s.stream>>>(h->d_small, h->d_local, s.K_ld, result);
            TMVB_CUDA(cudaGetLastError());
            s.st.kernel_launches++;
        }
    } else {
        lda_alpha_kernel<<<1, threads, 0, s.stream>>>(h->d_alpha64, h->d_alpha, h->d_small, h->d_local, (int)s.K, s.K_ld, (double)M_total, niter, ntol,
                                                      h->elbo_valid ? 1 : 0, result);
        TMVB_CUDA(cudaGetLastError());
    }
    s.st.kernel_launches++;
    h->alpha_on_device = true;
    h->elbo_dev_valid = h->elbo_valid;
    if (alpha_out) {
        TMVB_TRY(sync_alpha(h));
        for (int64_t i = 0; i < s.K; i++) alpha_out[i] = std::max((float)h->h_alpha[i], 1.1754944e-38f);
    }
    return 0;
}

int tmvb_lda_elbo(tmvb_lda_t h, int mode, int64_t M_total, double *elbo_docs, double *elbo_global)
{
    TMVB_CHECK_ARG(h && elbo_docs && elbo_global, "NULL argument");
    TMVB_CHECK_ARG(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    const int K = (int)s.K, K_ld = s.K_ld;
    if (mode == 0) {
        TMVB_CHECK_ARG(h->elbo_valid && h->elbo_dev_valid, "mode 0 needs estep(want_elbo=1), mstep, update_alpha in this order");
        // assembled by lda_alpha_kernel:  docs + M (lnG(sum alpha) - sum lnG(alpha))                      [Elogptheta, LDA.jl:51]
        //   + sum_i (alpha_i - alpha_estep_i - eps) Esum_i   [dot(alpha .- 1, Elogtheta) plus the (1 - alpha_estep) . Esum left over
        //                                                     from the per-document entropy/Elogpz terms, see lda_estep_kernel]
        //   + sum_ij S_ij ln(beta_ij + eps)                                                              [Elogpw over the statistics]
        TMVB_CUDA(cudaMemcpyAsync(s.h_pinned, h->d_local + 2 * K_ld + 1, 8, cudaMemcpyDeviceToHost, s.stream));
        // the read-back is the one host synchronisation of an iteration: the time-out flag of the peer exchange rides along, so
        // an exchange that gave up on a missing rank (and reduced half-written statistics) raises here instead of going unnoticed
        const bool peers = h->comm.connected && h->comm.d_ctl;
        if (peers) TMVB_CUDA(cudaMemcpyAsync(s.h_pinned + 1, h->comm.d_ctl + 128, 8, cudaMemcpyDeviceToHost, s.stream));
        TMVB_CUDA(cudaStreamSynchronize(s.stream));
        s.st.d2h_bytes += peers ? 16 : 8;
        if (peers) TMVB_TRY(comm_raise(h, reinterpret_cast<const unsigned *>(s.h_pinned + 1)[1]));
        *elbo_docs = s.h_pinned[0];
        *elbo_global = 0.0;
        (void)K;
        return 0;
    }
    TMVB_TRY(sync_alpha(h));
    TMVB_CHECK_ARG(s.corpus_set, "set_corpus has not been called");
    LdaDev p = dev_view(h);
    double *out = h->d_local + 2 * K_ld;
    TMVB_CUDA(cudaMemsetAsync(out, 0, 8, s.stream));
    if (s.M > 0) {
        const int grid = grid_for(s.M * 32, 128, s.n_sm);
        const double lga = lg_alpha_term(h->h_alpha);
        const float *bo = s.d_beta[s.cur ^ 1];
#define TMVB_ELBO_LAUNCH(T, R) lda_elbo_kernel<T, R><<<grid, 128, 0, s.stream>>>(p, bo, lga, out)
        if (mode == 2) {
            if (K <= 32) TMVB_ELBO_LAUNCH(double, 1); else if (K <= 64) TMVB_ELBO_LAUNCH(double, 2);
            else if (K <= 128) TMVB_ELBO_LAUNCH(double, 4); else TMVB_ELBO_LAUNCH(double, 8);
        } else if (env_int("TMVB_ELBO_LITERAL", 0)) {
            if (K <= 32) TMVB_ELBO_LAUNCH(float, 1); else if (K <= 64) TMVB_ELBO_LAUNCH(float, 2);
            else if (K <= 128) TMVB_ELBO_LAUNCH(float, 4); else TMVB_ELBO_LAUNCH(float, 8);
        } else if (h->beta_fresh && h->E_fresh && !env_int("TMVB_ELBO_NOFRESH", 0)) {
            // the lagged copies equal the current state (nothing ran since the upload): the one-dot-product-per-token form
            kLdaElboFresh[s.layout]<<<grid, 128, 0, s.stream>>>(p, lga, out);
        } else {
            const long long n = (long long)s.V * K_ld;
            TMVB_TRY(shard_scratch(&s, (size_t)std::max<long long>(n, 1) * 4));
            float *D = (float *)s.d_scratch;
            lda_logratio_kernel<<<grid_for(n, 256, s.n_sm), 256, 0, s.stream>>>(p.beta, bo, D, n, K, K_ld);
            s.st.kernel_launches++;
#define TMVB_ELBO_FAST(R) lda_elbo_fast_kernel<R><<<grid, 128, 0, s.stream>>>(p, bo, D, lga, out)
            if (K <= 32) TMVB_ELBO_FAST(1); else if (K <= 64) TMVB_ELBO_FAST(2);
            else if (K <= 128) TMVB_ELBO_FAST(4); else TMVB_ELBO_FAST(8);
#undef TMVB_ELBO_FAST
        }
#undef TMVB_ELBO_LAUNCH
        TMVB_CUDA(cudaGetLastError());
        s.st.kernel_launches++;
    }
    TMVB_CUDA(cudaMemcpyAsync(s.h_pinned, out, 8, cudaMemcpyDeviceToHost, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    s.st.d2h_bytes += 8;
    *elbo_docs = s.h_pinned[0];
    *elbo_global = 0.0;
    return 0;
}

int tmvb_lda_arm_host_mirror(tmvb_lda_t h, float *Elogtheta, float *gamma)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    h->mirror_armed = false;
    if (!Elogtheta && !gamma) return 0;   // disarm
    TMVB_CHECK_ARG(Elogtheta && gamma, "the host mirror needs both arrays");
    TMVB_CHECK_ARG(s.corpus_set, "set_corpus has not been called");
    long long covered = 0;
    for (const Bucket &b : s.buckets) covered += b.doc_end - b.doc_begin;
    TMVB_CHECK_ARG(covered == (long long)s.M, "internal: the E-step launches do not cover every document");
    float *dev[2] = {nullptr, nullptr};
    float *host[2] = {Elogtheta, gamma};
    for (int a = 0; a < 2; a++) {
        cudaPointerAttributes at;
        memset(&at, 0, sizeof(at));
        if (cudaPointerGetAttributes(&at, host[a]) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
            cudaGetLastError();
            return fail(-1, "invalid argument: the host mirror arrays must be page-locked (cudaHostAlloc / cudaHostRegister) and mapped");
        }
        dev[a] = (float *)at.devicePointer;
    }
    h->mirror_E = dev[0];
    h->mirror_gamma = dev[1];
    h->mirror_E_host = Elogtheta;
    h->mirror_gamma_host = gamma;
    h->mirror_armed = true;
    return 0;
}

int tmvb_lda_download(tmvb_lda_t h, float *alpha, float *beta, float *Elogtheta, float *gamma)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    if (h->comm.connected && h->comm.d_ctl) {   // update_host! of a run that never read an ELBO (checkelbo = Inf)
        int st = 0;
        TMVB_TRY(tmvb_lda_comm_status(h, &st));
    }
    if (alpha) {
        TMVB_TRY(sync_alpha(h));
        for (int64_t i = 0; i < s.K; i++) alpha[i] = std::max((float)h->h_alpha[i], 1.1754944e-38f);
    }
    TMVB_TRY(shard_download_rows(&s, s.d_beta[s.cur], beta, s.V, nullptr));
    if (s.M > 0 && (Elogtheta || gamma)) TMVB_CHECK_ARG(s.corpus_set, "set_corpus has not been called");
    // rows the last E-step already wrote into these very arrays (host mirror) are complete once the stream is idle
    const bool have_E = h->mirror_valid && Elogtheta == h->mirror_E_host, have_g = h->mirror_valid && gamma == h->mirror_gamma_host;
    if (have_E || have_g) TMVB_CUDA(cudaStreamSynchronize(s.stream));
    if (!have_E) TMVB_TRY(shard_download_rows(&s, h->d_Elogtheta, Elogtheta, s.M, s.d_perm));
    if (!have_g) TMVB_TRY(shard_download_rows(&s, h->d_gamma, gamma, s.M, s.d_perm));
    return 0;
}

int tmvb_lda_download_old(tmvb_lda_t h, float *beta_old, float *Elogtheta_old)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    TMVB_TRY(shard_download_rows(&s, s.d_beta[s.cur ^ 1], beta_old, s.V, nullptr));
    TMVB_TRY(shard_download_rows(&s, h->d_Elogtheta_old, Elogtheta_old, s.M, s.d_perm));
    return 0;
}

int tmvb_lda_materialize_phi(tmvb_lda_t h, float *phi)
{
    TMVB_CHECK_ARG(h && phi, "NULL argument");
    Shard &s = h->s;
    TMVB_CHECK_ARG(s.corpus_set, "set_corpus has not been called");
    TMVB_CUDA(cudaSetDevice(s.device));
    if (s.nnz == 0) return 0;
    const size_t bytes = (size_t)s.nnz * s.K * 4;
    TMVB_TRY(shard_scratch(&s, bytes));
    LdaDev p = dev_view(h);
    lda_phi_kernel<<<grid_for(s.M * 32, 128, s.n_sm), 128, 0, s.stream>>>(p, s.d_beta[s.cur ^ 1], s.d_src_off, (float *)s.d_scratch);
    TMVB_CUDA(cudaGetLastError());
    s.st.kernel_launches++;
    TMVB_CUDA(cudaMemcpyAsync(phi, s.d_scratch, bytes, cudaMemcpyDeviceToHost, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    s.st.d2h_bytes += bytes;
    return 0;
}

int tmvb_lda_topics(tmvb_lda_t h, int32_t *topics)
{
    TMVB_CHECK_ARG(h && topics, "NULL argument");
    TMVB_CUDA(cudaSetDevice(h->s.device));
    return shard_topics(&h->s, h->s.d_beta[h->s.cur], nullptr, topics);
}

int tmvb_lda_sync(tmvb_lda_t h)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CUDA(cudaSetDevice(h->s.device));
    TMVB_CUDA(cudaStreamSynchronize(h->s.stream));
    return 0;
}

int tmvb_lda_get_stats(tmvb_lda_t h, tmvb_stats *out)
{
    TMVB_CHECK_ARG(h && out, "NULL argument");
    return shard_get_stats(&h->s, h->d_small + h->s.K_ld + 1, out);
}

}  // extern "C"
