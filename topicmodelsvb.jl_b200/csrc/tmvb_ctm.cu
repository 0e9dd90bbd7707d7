// tmvb_ctm.cu -- correlated topic model (logistic-normal) coordinate-ascent VB on sm_100a and the tmvb_ctm_* C ABI.
//
// Reference semantics: the CPU model src/CTM.jl (update order phi -> logzeta -> vsq -> lambda per sweep, per-document
// stopping rule ||lambda - lambda_old|| < vtol, true log-sum-exp, lagged-phi ELBO).  Replaced: src/gpuCTM.jl's 9 OpenCL
// kernels + `linsolve` (utils.jl:60-90) + modelutils.jl:400-435,518-537.
//
// One warp per document (tmvb_estep.cuh).  Per sweep:
//   token phase  phi_n = softmax_i(log beta[i,w_n] + lambda_i) == beta e / (beta . e), e = exp(lambda - max lambda)
//                (CTM.jl:175-178) -- the same two FMA passes over the shared-memory tile as LDA, no epsilon
//   K phase      logzeta = logsumexp(lambda + vsq/2)                                   (CTM.jl:169-171)
//                vsq: per-coordinate Newton with back-tracking, one coordinate per lane  (CTM.jl:146-165)
//                lambda: Newton, (invsigma + C diag(w)) x = grad solved by an in-warp Cholesky factorisation in
//                shared memory (lane l owns rows l, l+32; LDS.128 dot products) replacing the reference's
//                Gauss-Jordan `linsolve`                                               (CTM.jl:129-142)
// M-step: beta as LDA; sigma = (diag sum vsq + sum (lambda-mu)(lambda-mu)')/M, invsigma, mu in fp64 on the host from the
// second moments sum_d lambda_d lambda_d' accumulated on the device (CTM.jl:102-111; update_sigma! uses the old mu).
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "tmvb_comm.cuh"
#include "tmvb_filt.cuh"
#include "tmvb_shard.cuh"

namespace tmvb {

struct CtmDev {
    int K, K_ld, V, RS, KP;
    long long M;
    const float *beta;
    float *stats;
    const long long *doc_off;
    const int *terms;
    const float *counts;
    float *lambda, *lambda_old, *vsq, *logzeta;
    const float *mu;        // [K_ld]
    const float *invsigma;  // [K][KP] row-major, KP = 4 (mod 8)
    double *small;          // [0] per-document ELBO terms, [1] sweep counter
    // filtered CTM (fCTM.jl) only: log2 table of beta, (1 - eta) kappa | eta, update_kappa! statistics, per-token tau / tau_old
    const float *L, *kq;
    float *kstats, *tau, *tau_old;
    int niter;
    float ntol;
    int viter;
    float vtol;
    int stage_bulk, dbg;
};

static int ctm_kp(int K)
{
    int kp = K;
    while (kp % 8 != 4) kp++;
    return kp;
}
// shared memory beyond the tile: mbarrier | gs [S][RS] | e_s [RS] | invsigma [K][KP] | L [K][KP] | vec [K_ld] | dinv [K_ld]
// (invsigma stays in global memory -- read-only, shared by every CTA, L1/L2 resident: the kernel is bound by the latency of the
// in-warp Cholesky chains, so what counts is how many documents an SM holds, i.e. shared memory per warp)
// the register factorisation (K_ld <= 32) borrows L for its 32 x 33 row store and vec | dinv for its two 32-float column buffers
__host__ __device__ constexpr int ctm_l_floats(int K, int KP, int K_ld) { return (K_ld <= 32 && K * KP < 32 * 33) ? 32 * 33 : K * KP; }
__host__ __device__ constexpr int ctm_vd_floats(int K_ld) { return 2 * K_ld < 64 ? 64 : 2 * K_ld; }
static size_t ctm_fixed_smem(int RS, int lpt, int K, int K_ld)
{
    return 16 + (size_t)(32 / lpt) * RS * 4 + (size_t)RS * 4 + (size_t)ctm_l_floats(K, ctm_kp(K), K_ld) * 4 + (size_t)ctm_vd_floats(K_ld) * 4;
}


// entry j of a vector distributed over the warp as v[r] on lane l for j = l + 32 r
template <int R>
__device__ __forceinline__ float warp_entry(const float (&v)[R], int j)
{
    float out = 0.0f;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float x = __shfl_sync(0xffffffffu, v[r], j & 31);
        if ((j >> 5) == r) out = x;
    }
    return out;
}

// Cholesky A = L L' of the SPD matrix A = inv_s + diag(w) (lane l owns rows l + 32 r), left-looking by columns,
// dot products over 16-byte chunks; L is zero-initialised so that partial chunks read zeros.  dinv[j] = 1 / L[j][j].
template <int R>
__device__ __forceinline__ void warp_cholesky(const float *inv_s, float *L_s, float *dinv_s, const float (&w)[R], int K, int KP, int lane)
{
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = lane; q < (K * KP) >> 2; q += 32) reinterpret_cast<float4 *>(L_s)[q] = z4;
    __syncwarp();
    for (int j = 0; j < K; j++) {
        const int nch = (j + 3) >> 2;
        const float4 *rowj = reinterpret_cast<const float4 *>(L_s + j * KP);
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            acc[r] = 0.0f;
            if (i >= j && i < K) {
                float a0 = inv_s[i * KP + j] + ((i == j) ? w[r] : 0.0f), a1 = 0.0f;
                const float4 *rowi = reinterpret_cast<const float4 *>(L_s + i * KP);
                for (int c = 0; c < nch; c++) {
                    const float4 a = rowi[c], b = rowj[c];
                    a0 = fmaf(-a.x, b.x, a0);
                    a1 = fmaf(-a.y, b.y, a1);
                    a0 = fmaf(-a.z, b.z, a0);
                    a1 = fmaf(-a.w, b.w, a1);
                }
                acc[r] = a0 + a1;
            }
        }
        const float djj = warp_entry<R>(acc, j);
        const float di = rsqrtf(fmaxf(djj, 1e-30f));
        __syncwarp();  // every lane has finished reading row j before column j is written into it
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            if (i >= j && i < K) L_s[i * KP + j] = acc[r] * di;  // i == j: djj / sqrt(djj) = sqrt(djj)
        }
        if (lane == 0) dinv_s[j] = di;
        __syncwarp();
    }
}

// x = (L L')^-1 b, b and x distributed like the rows (lane l holds entries l + 32 r)
template <int R>
__device__ __forceinline__ void warp_chol_solve(const float *L_s, const float *dinv_s, float (&b)[R], int K, int KP, int lane)
{
    for (int j = 0; j < K; j++) {
        const float bj = warp_entry<R>(b, j);
        const float yj = bj * dinv_s[j];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            if (i > j && i < K) b[r] = fmaf(-L_s[i * KP + j], yj, b[r]);
            if (i == j) b[r] = yj;
        }
    }
    for (int j = K - 1; j >= 0; j--) {
        const float bj = warp_entry<R>(b, j);
        const float xj = bj * dinv_s[j];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            if (i < j) b[r] = fmaf(-L_s[j * KP + i], xj, b[r]);
            if (i == j) b[r] = xj;
        }
    }
}

// ---- K_ld <= 32: the whole factorisation in registers ------------------------------------------------------------------
// Lane i keeps row i of the lower triangle (a[0..i]); right-looking Cholesky with the loops fully unrolled (KP = K_ld is a
// compile-time constant of the lane layout), the column broadcast by shuffles: KP (KP - 1) / 2 x (SHFL + FFMA), no shared
// memory, no loop or branch instructions.  The shared-memory version above spends ~2 900 warp-instructions per factorisation at
// K = 30 (450 of them FFMA: chunk loops, address arithmetic, two warp barriers per column) and was 65 % of the E-step; this one
// ~1 100.  Rows / columns beyond K are the identity.  dinv = 1 / l_ii of this lane's row.
#ifndef TMVB_CTM_REGCHOL
#define TMVB_CTM_REGCHOL 1   // 0: the shared-memory factorisation for every K (A/B switch)
#endif
template <int KP>
__device__ __forceinline__ void reg_chol_load(const float *__restrict__ inv_s, int ldp, float w, int K, int lane, float (&a)[KP])
{
    if (lane < K) {
        const float4 *row = reinterpret_cast<const float4 *>(inv_s + (size_t)lane * ldp);
#pragma unroll
        for (int c = 0; c < KP / 4; c++) {
            const float4 v = (4 * c < K) ? __ldg(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);   // pad columns of inv_s (j >= K) are zero
            a[4 * c] = v.x;
            a[4 * c + 1] = v.y;
            a[4 * c + 2] = v.z;
            a[4 * c + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < KP; k++) a[k] = 0.0f;
    }
#pragma unroll
    for (int k = 0; k < KP; k++)
        if (k == lane) a[k] = (lane < K) ? a[k] + w : 1.0f;
}
// col_s: [2][32] floats of warp-private shared memory (column j of L is published there, double-buffered by the parity of j, and read
// back as broadcast LDS.128: 32 shuffles + ~124 LDS.128 per factorisation instead of 496 shuffles -- the shuffle pipe was the limit)
template <int KP>
__device__ __forceinline__ float reg_cholesky(float (&a)[KP], int lane, float *col_s)
{
    float dinv = 1.0f;
#pragma unroll
    for (int j = 0; j < KP; j++) {
        const float djj = __shfl_sync(0xffffffffu, a[j], j);
        const float di = rsqrtf(fmaxf(djj, 1e-30f));
        a[j] *= di;                       // l_ij for the lanes i >= j (lane j: sqrt(d_jj)); unused garbage above the diagonal
        if (lane == j) dinv = di;
        if (j + 1 < KP) {
            float *cs = col_s + 32 * (j & 1);
            cs[lane] = a[j];
            __syncwarp();
#pragma unroll
            for (int c = (j + 1) / 4; c < KP / 4; c++) {
                const float4 v = reinterpret_cast<const float4 *>(cs)[c];
                if (4 * c + 0 > j) a[4 * c + 0] = fmaf(-a[j], v.x, a[4 * c + 0]);   // a_ik -= l_ij l_kj  (meaningful for i >= k)
                if (4 * c + 1 > j) a[4 * c + 1] = fmaf(-a[j], v.y, a[4 * c + 1]);
                if (4 * c + 2 > j) a[4 * c + 2] = fmaf(-a[j], v.z, a[4 * c + 2]);
                if (4 * c + 3 > j) a[4 * c + 3] = fmaf(-a[j], v.w, a[4 * c + 3]);
            }
        }
    }
    __syncwarp();
    return dinv;
}
// x = (L L')^-1 b, b distributed one entry per lane; L row-wise in registers as left by reg_cholesky.  Lt_s: [KP][33] floats of
// warp-private shared memory: the rows of L are stored there once so that the back substitution reads column j of L' (= row
// entries l_ij of the lanes i > j) as one conflict-free LDS per step instead of a five-step warp reduction.
template <int KP>
__device__ __forceinline__ float reg_chol_solve(const float (&a)[KP], float dinv, float b, int lane, float *Lt_s)
{
    if (lane < KP) {
#pragma unroll
        for (int k = 0; k < KP; k++) Lt_s[k * 33 + lane] = a[k];   // Lt_s[k][i] = l_ik: lanes write consecutive words
    }
#pragma unroll
    for (int j = 0; j < KP; j++) {        // L y = b: lane j finishes y_j and broadcasts it
        const float yj = __shfl_sync(0xffffffffu, b * dinv, j);
        if (lane > j) b = fmaf(-a[j], yj, b);
        if (lane == j) b = yj;
    }
    __syncwarp();
#pragma unroll
    for (int i = KP - 1; i >= 0; i--) {   // L' x = y: lane i finishes x_i and broadcasts it; lane j < i subtracts l_ij x_i
        const float xi = __shfl_sync(0xffffffffu, b * dinv, i);
        // l_ij of row i lives in lane i; lane j reads it from the row store: Lt_s[j][i] would be the transposed view of lane j's
        // own registers, so read the copy lane i wrote: element (k = j, lane = i)
        if (lane < i) b = fmaf(-Lt_s[lane * 33 + i], xi, b);
        if (lane == i) b = xi;
    }
    __syncwarp();
    return b;
}

#ifndef TMVB_CTM_MIN_CTAS
#define TMVB_CTM_MIN_CTAS 16   // 128 registers per thread
#endif
template <int LPT, int CPL, bool ELBO>
__global__ void __launch_bounds__(32, TMVB_CTM_MIN_CTAS) ctm_estep_kernel(const CtmDev p, int doc_begin, int doc_end, int cap, int cap2, int *counter)
{
    constexpr int S = 32 / LPT;
    constexpr int R = (LPT * CPL + 7) / 8;
    static_assert(R <= 4, "CTM supports K <= 128 (the Cholesky factor lives in shared memory)");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int kl = lane % LPT, ts = lane / LPT;
    const int K = p.K, K_ld = p.K_ld, CH = K_ld >> 2, RS = p.RS, KP = p.KP;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    (void)cap2;

    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw);
    float *gs = reinterpret_cast<float *>(smem_raw + 16);  // [S][RS]
    float *e_s = gs + (size_t)S * RS;                      // [RS]
    float *L_s = e_s + RS;                                 // [K][KP]
    const float *inv_s = p.invsigma;                       // global: L1-resident, shared by all CTAs
    float *vec_s = L_s + ctm_l_floats(K, KP, K_ld);        // [K_ld]
    float *dinv_s = vec_s + K_ld;                          // [K_ld]  (vec | dinv: at least 64 floats)
    float *tile = vec_s + ctm_vd_floats(K_ld);             // [cap][RS]
    float *cnt_s = tile + (size_t)cap * RS;
    int *term_s = reinterpret_cast<int *>(cnt_s + cap);

    float mu_k[R], isd_k[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = lane + 32 * r;
        mu_k[r] = (i < K) ? p.mu[i] : 0.0f;
        isd_k[r] = (i < K) ? p.invsigma[i * KP + i] : 1.0f;
    }
    double elbo_thr = 0.0;
    unsigned long long sweeps_thr = 0;
    unsigned phase = 0;
    if (p.stage_bulk) {
        if (lane == 0) mbar_init(mbar, 1);
    }
    __syncwarp();

    for (;;) {
        int d = 0;
        if (lane == 0) d = doc_begin + atomicAdd(counter, 1);
        d = __shfl_sync(0xffffffffu, d, 0);
        if (d >= doc_end) break;
        const long long o = p.doc_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o);
        const int ns = min(Nd, cap);
        const bool ovf = Nd > cap;

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
        stage_rows(tile, term_s, p.beta, ns, K_ld, RS, lane, mbar, p.stage_bulk);
        float lam_k[R], lold_k[R], vsq_k[R], phic_k[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            lam_k[r] = (i < K) ? p.lambda[(size_t)d * K_ld + i] : 0.0f;
            vsq_k[r] = (i < K) ? p.vsq[(size_t)d * K_ld + i] : 1.0f;
            lold_k[r] = lam_k[r];
            phic_k[r] = 0.0f;
        }
        const float Cd = warp_sum(csum);
        stage_wait(mbar, phase, p.stage_bulk);

        TokArgs ta;
        ta.tile = tile;
        ta.cnt_s = cnt_s;
        ta.term_s = term_s;
        ta.gtable = p.beta;
        ta.gterms = p.terms + o;
        ta.gcounts = p.counts + o;
        ta.stats = p.stats;
        ta.Nd = Nd;
        ta.cap = cap;
        ta.rounds = (Nd + S - 1) / S;
        ta.K = K;
        ta.K_ld = K_ld;
        ta.RS = RS;
        ta.dbg = p.dbg;
        ta.r0 = 0;
        ta.rstep = 1;

        float4 e[CPL];
        float logzeta = 0.0f;
        int v = 0;
        for (;;) {
            // ---- update_phi! (CTM.jl:175-178): e = exp(lambda - max lambda), phi*counts = e .* g
            float lmax = -INFINITY;
#pragma unroll
            for (int r = 0; r < R; r++)
                if (lane + 32 * r < K) lmax = fmaxf(lmax, lam_k[r]);
            lmax = warp_max(lmax);
            float e_k[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = lane + 32 * r;
                e_k[r] = (i < K) ? __expf(lam_k[r] - lmax) : 0.0f;
                if (i < K_ld) e_s[i] = e_k[r];
            }
            __syncwarp();
#pragma unroll
            for (int m = 0; m < CPL; m++)
                e[m] = (m < CPL - 1 || kl + LPT * m < CH) ? reinterpret_cast<const float4 *>(e_s)[kl + LPT * m] : zero4;
            float4 g[CPL];
            float tsum = 0.0f;
#pragma unroll
            for (int m = 0; m < CPL; m++) g[m] = zero4;
            if (!ovf)
                tok_sweep<LPT, CPL, false, false>(ta, ts, kl, e, g, tsum);
            else
                tok_sweep<LPT, CPL, true, false>(ta, ts, kl, e, g, tsum);
#pragma unroll
            for (int m = 0; m < CPL; m++)
                if (m < CPL - 1 || kl + LPT * m < CH) reinterpret_cast<float4 *>(gs + ts * RS)[kl + LPT * m] = g[m];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = lane + 32 * r;
                phic_k[r] = (i < K) ? e_k[r] * owner_sum<S>(gs, RS, i) : 0.0f;
            }

            // ---- update_logzeta! (CTM.jl:169-171)
            float zmax = -INFINITY;
#pragma unroll
            for (int r = 0; r < R; r++)
                if (lane + 32 * r < K) zmax = fmaxf(zmax, fmaf(0.5f, vsq_k[r], lam_k[r]));
            zmax = warp_max(zmax);
            float zs = 0.0f;
#pragma unroll
            for (int r = 0; r < R; r++)
                if (lane + 32 * r < K) zs += __expf(fmaf(0.5f, vsq_k[r], lam_k[r]) - zmax);
            logzeta = zmax + logf(warp_sum(zs));

            // ---- update_vsq! (CTM.jl:146-165): Newton + back-tracking, one coordinate per lane slot
#pragma unroll
            for (int r = 0; r < R; r++) {
                bool active = lane + 32 * r < K;
                for (int it = 0; it < p.niter; it++) {
                    if (active) {
                        float rho = 1.0f;
                        const float ex = Cd * __expf(fmaf(0.5f, vsq_k[r], lam_k[r]) - logzeta);
                        const float grad = -0.5f * (isd_k[r] + ex - 1.0f / vsq_k[r]);
                        const float invhess = -1.0f / (0.25f * ex + 0.5f / (vsq_k[r] * vsq_k[r]));
                        const float pp = invhess * grad;
                        while (vsq_k[r] - rho * pp <= 0.0f) rho *= 0.5f;
                        vsq_k[r] -= rho * pp;
                        if (rho * fabsf(grad) < p.ntol) active = false;
                    }
                    if (!__any_sync(0xffffffffu, active)) break;
                }
                vsq_k[r] += TMVB_EPS;  // @positive model.vsq[d]
            }

            // ---- update_lambda! (CTM.jl:129-142)
#pragma unroll
            for (int r = 0; r < R; r++) lold_k[r] = lam_k[r];
            for (int it = 0; it < p.niter; it++) {
                float w_k[R], grad_k[R];
                __syncwarp();
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int i = lane + 32 * r;
                    w_k[r] = (i < K) ? Cd * __expf(fmaf(0.5f, vsq_k[r], lam_k[r]) - logzeta) : 0.0f;
                    if (i < K_ld) vec_s[i] = (i < K) ? mu_k[r] - lam_k[r] : 0.0f;
                }
                __syncwarp();
                float gn = 0.0f;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int i = lane + 32 * r;
                    grad_k[r] = 0.0f;
                    if (i < K) {
                        float a0 = 0.0f, a1 = 0.0f;
                        const float4 *row = reinterpret_cast<const float4 *>(inv_s + i * KP);
                        const float4 *vv = reinterpret_cast<const float4 *>(vec_s);
                        for (int c = 0; c < (K + 3) >> 2; c++) {  // inv_s pad columns (j >= K) are zero
                            const float4 a = row[c], b = vv[c];
                            a0 = fmaf(a.x, b.x, a0);
                            a1 = fmaf(a.y, b.y, a1);
                            a0 = fmaf(a.z, b.z, a0);
                            a1 = fmaf(a.w, b.w, a1);
                        }
                        grad_k[r] = (a0 + a1) + phic_k[r] - w_k[r];
                        gn = fmaf(grad_k[r], grad_k[r], gn);
                    }
                }
                gn = warp_sum(gn);
                if (TMVB_CTM_REGCHOL && 4 * LPT * CPL <= 32 && K_ld == 4 * LPT * CPL) {
                    constexpr int KC = (4 * LPT * CPL <= 32) ? 4 * LPT * CPL : 4;   // (the second arm keeps the dead instantiation small)
                    float a_row[KC];
                    reg_chol_load<KC>(inv_s, KP, w_k[0], K, lane, a_row);
                    const float di = reg_cholesky<KC>(a_row, lane, vec_s);           // vec | dinv: >= 64 floats (the gradient's vec_s is dead by now)
                    grad_k[0] = reg_chol_solve<KC>(a_row, di, grad_k[0], lane, L_s);   // L_s: >= 32 * 33 floats (ctm_l_floats)
                } else {
                    warp_cholesky<R>(inv_s, L_s, dinv_s, w_k, K, KP, lane);
                    warp_chol_solve<R>(L_s, dinv_s, grad_k, K, KP, lane);
                }
#pragma unroll
                for (int r = 0; r < R; r++)
                    if (lane + 32 * r < K) lam_k[r] += grad_k[r];
                if (sqrtf(gn) < p.ntol) break;
            }
            float dl = 0.0f;
#pragma unroll
            for (int r = 0; r < R; r++)
                if (lane + 32 * r < K) dl = fmaf(lam_k[r] - lold_k[r], lam_k[r] - lold_k[r], dl);
            dl = warp_sum(dl);
            v++;
            if (sqrtf(dl) < p.vtol || v >= p.viter) break;  // CTM.jl:200
        }

        // update_beta!(model, d) (CTM.jl:122-125): scatter the last phi (built from lambda_old)
        float ent = 0.0f;
        if (!ovf)
            tok_final<LPT, CPL, false, false, ELBO>(ta, ts, kl, e, ent);
        else
            tok_final<LPT, CPL, true, false, ELBO>(ta, ts, kl, e, ent);

        float a = 0.0f, se = 0.0f;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            if (i < K_ld) {
                const bool ok = i < K;
                p.lambda[(size_t)d * K_ld + i] = ok ? lam_k[r] : 0.0f;
                p.lambda_old[(size_t)d * K_ld + i] = ok ? lold_k[r] : 0.0f;
                p.vsq[(size_t)d * K_ld + i] = ok ? vsq_k[r] : 0.0f;
                if (ELBO && ok) {
                    a = fmaf(phic_k[r], lam_k[r], a);                                   // Elogpz first term, CTM.jl:64
                    a += 0.5f * logf(vsq_k[r]);                                          // entropy(MvNormal) part, CTM.jl:77
                    se += __expf(fmaf(0.5f, vsq_k[r], lam_k[r]) - logzeta);
                }
            }
        }
        if (lane == 0) p.logzeta[d] = logzeta;
        if (ELBO) {
            se = warp_sum(se);
            if (lane == 0) a -= Cd * (se + logzeta - 1.0f);                              // CTM.jl:64
            elbo_thr += (double)a + (double)ent;
        }
        if (lane == 0) sweeps_thr += (unsigned long long)v;
    }
    if (ELBO) {
        const double tot = warp_sum_d(elbo_thr);
        if (lane == 0 && tot != 0.0) atomicAdd(p.small, tot);
    }
    if (lane == 0 && sweeps_thr) atomicAdd(p.small + 1, (double)sweeps_thr);
}

// ------------------------------------------------------------------ filtered CTM (src/fCTM.jl) ----------
// The inner loop of train!(::fCTM) (fCTM.jl:258-268): update_phi!, update_tau!, update_logzeta!, update_lambda!, update_vsq! -- in
// THAT order (lambda before vsq, unlike CTM.jl:196-199) -- with the filtered token pass of tmvb_filt.cuh: phi = softmax_i(tau_n
// log2(beta + eps) + (lambda_i - max lambda) log2 e) from the log2 table p.L, update_tau! fused, tau / tau_old in global memory
// (L2-resident between sweeps; the tile holds 16 rows only, as in ctm_estep_kernel).  The Newton iterations and the Cholesky
// solve are those of ctm_estep_kernel.  Scatter: phi tau c into the statistics (fCTM.jl:175-178), (1 - tau) c into kstats
// (fCTM.jl:162-165).  The ELBO comes from fctm_elbo_kernel (no fused partials).
template <int LPT, int CPL>
__global__ void __launch_bounds__(32, 8) fctm_estep_kernel(const CtmDev p, int doc_begin, int doc_end, int cap, int cap2, int *counter)
{
    constexpr int S = 32 / LPT;
    constexpr int R = (LPT * CPL + 7) / 8;
    static_assert(R <= 4, "fCTM supports K <= 128");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int kl = lane % LPT, ts = lane / LPT;
    const int K = p.K, K_ld = p.K_ld, CH = K_ld >> 2, RS = p.RS, KP = p.KP;
    (void)cap2;
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw);
    float *gs = reinterpret_cast<float *>(smem_raw + 16);  // [S][RS]
    float *e_s = gs + (size_t)S * RS;                      // [RS]
    float *L_s = e_s + RS;                                 // [K][KP]
    const float *inv_s = p.invsigma;
    float *vec_s = L_s + ctm_l_floats(K, KP, K_ld);        // [K_ld]
    float *dinv_s = vec_s + K_ld;                          // [K_ld]  (vec | dinv: at least 64 floats)
    float *tile = vec_s + ctm_vd_floats(K_ld);             // [cap][RS]
    float *cnt_s = tile + (size_t)cap * RS;
    int *term_s = reinterpret_cast<int *>(cnt_s + cap);

    float mu_k[R], isd_k[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = lane + 32 * r;
        mu_k[r] = (i < K) ? p.mu[i] : 0.0f;
        isd_k[r] = (i < K) ? p.invsigma[i * KP + i] : 1.0f;
    }
    const float eta = __ldg(p.kq + p.V);
    unsigned long long sweeps_thr = 0;
    unsigned phase = 0;
    if (lane == 0) mbar_init(mbar, 1);
    __syncwarp();

    for (;;) {
        int d = 0;
        if (lane == 0) d = doc_begin + atomicAdd(counter, 1);
        d = __shfl_sync(0xffffffffu, d, 0);
        if (d >= doc_end) break;
        const long long o = p.doc_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o);
        const int ns = min(Nd, cap);
        const int rounds = (Nd + S - 1) / S;

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
        stage_rows(tile, term_s, p.L, ns, K_ld, RS, lane, mbar, 1);
        float lam_k[R], lold_k[R], vsq_k[R], phic_k[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            lam_k[r] = (i < K) ? p.lambda[(size_t)d * K_ld + i] : 0.0f;
            vsq_k[r] = (i < K) ? p.vsq[(size_t)d * K_ld + i] : 1.0f;
            lold_k[r] = lam_k[r];
            phic_k[r] = 0.0f;
        }
        const float Cd = warp_sum(csum);
        stage_wait(mbar, phase, 1);

        float logzeta = 0.0f;
        int v = 0;
        for (;;) {
            // ---- update_phi! (fCTM.jl:239-242) + update_tau! (fCTM.jl:230-235)
            float lmax = -INFINITY;
#pragma unroll
            for (int r = 0; r < R; r++)
                if (lane + 32 * r < K) lmax = fmaxf(lmax, lam_k[r]);
            lmax = warp_max(lmax);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = lane + 32 * r;
                if (i < K_ld) e_s[i] = (i < K) ? (lam_k[r] - lmax) * kLog2e : kPadE;
            }
            __syncwarp();
            f32x2 E01[CPL], E23[CPL], g01[CPL], g23[CPL];
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                const bool in = kl + LPT * m < CH;
                const float4 E = in ? reinterpret_cast<const float4 *>(e_s)[kl + LPT * m] : make_float4(kPadE, kPadE, kPadE, kPadE);
                E01[m] = pk2(E.x, E.y);
                E23[m] = pk2(E.z, E.w);
                g01[m] = g23[m] = 0ull;
            }
            for (int r = 0; r < rounds; r++) {
                const int n = r * S + ts;
                const bool ok = n < Nd;
                const int nn = ok ? n : 0;
                const bool in_tile = nn < cap;
                const int term = in_tile ? term_s[nn] : __ldg(p.terms + o + nn);
                ulonglong2 b[CPL];
                flda_load_row<LPT, CPL>(tile, p.L, RS, K_ld, CH, nn, cap, term, kl, b);
                const float c = ok ? (in_tile ? cnt_s[nn] : __ldg(p.counts + o + nn)) : 0.0f;
                const float tau = __ldcg(p.tau + o + nn);
                f32x2 p01[CPL], p23[CPL];
                float sn, q;
                flda_row<CPL>(b, E01, E23, tau, p01, p23, sn, q);
                sn = group_sum<LPT>(sn);
                q = group_sum<LPT>(q);
                const float rs = rcp_ftz(sn);
                const float t = c * rs;
                const f32x2 t2 = pk2(t, t);
#pragma unroll
                for (int m = 0; m < CPL; m++) {
                    g01[m] = fma2(p01[m], t2, g01[m]);
                    g23[m] = fma2(p23[m], t2, g23[m]);
                }
                __syncwarp();   // every lane of the token has read tau before lane kl = 0 replaces it
                if (ok && kl == 0) {
                    const float den = (eta + __ldg(p.kq + term) * ex2_ftz(-q * rs)) + TMVB_EPS;
                    __stcg(p.tau_old + o + nn, tau);
                    __stcg(p.tau + o + nn, fast_div_pos(eta, den));
                }
            }
#pragma unroll
            for (int m = 0; m < CPL; m++)
                if (kl + LPT * m < CH) {
                    float4 gv;
                    unpk2(g01[m], gv.x, gv.y);
                    unpk2(g23[m], gv.z, gv.w);
                    reinterpret_cast<float4 *>(gs + (size_t)ts * RS)[kl + LPT * m] = gv;
                }
            __syncwarp();
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = lane + 32 * r;
                phic_k[r] = (i < K) ? owner_sum<S>(gs, RS, i) : 0.0f;
            }

            // ---- update_logzeta! (fCTM.jl:224-226)
            float zmax = -INFINITY;
#pragma unroll
            for (int r = 0; r < R; r++)
                if (lane + 32 * r < K) zmax = fmaxf(zmax, fmaf(0.5f, vsq_k[r], lam_k[r]));
            zmax = warp_max(zmax);
            float zs = 0.0f;
#pragma unroll
            for (int r = 0; r < R; r++)
                if (lane + 32 * r < K) zs += __expf(fmaf(0.5f, vsq_k[r], lam_k[r]) - zmax);
            logzeta = zmax + logf(warp_sum(zs));

            // ---- update_lambda! (fCTM.jl:181-196)
#pragma unroll
            for (int r = 0; r < R; r++) lold_k[r] = lam_k[r];
            for (int it = 0; it < p.niter; it++) {
                float w_k[R], grad_k[R];
                __syncwarp();
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int i = lane + 32 * r;
                    w_k[r] = (i < K) ? Cd * __expf(fmaf(0.5f, vsq_k[r], lam_k[r]) - logzeta) : 0.0f;
                    if (i < K_ld) vec_s[i] = (i < K) ? mu_k[r] - lam_k[r] : 0.0f;
                }
                __syncwarp();
                float gn = 0.0f;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int i = lane + 32 * r;
                    grad_k[r] = 0.0f;
                    if (i < K) {
                        float a0 = 0.0f, a1 = 0.0f;
                        const float4 *row = reinterpret_cast<const float4 *>(inv_s + i * KP);
                        const float4 *vv = reinterpret_cast<const float4 *>(vec_s);
                        for (int c = 0; c < (K + 3) >> 2; c++) {
                            const float4 a = row[c], b = vv[c];
                            a0 = fmaf(a.x, b.x, a0);
                            a1 = fmaf(a.y, b.y, a1);
                            a0 = fmaf(a.z, b.z, a0);
                            a1 = fmaf(a.w, b.w, a1);
                        }
                        grad_k[r] = (a0 + a1) + phic_k[r] - w_k[r];
                        gn = fmaf(grad_k[r], grad_k[r], gn);
                    }
                }
                gn = warp_sum(gn);
                if (TMVB_CTM_REGCHOL && 4 * LPT * CPL <= 32 && K_ld == 4 * LPT * CPL) {
                    constexpr int KC = (4 * LPT * CPL <= 32) ? 4 * LPT * CPL : 4;   // (the second arm keeps the dead instantiation small)
                    float a_row[KC];
                    reg_chol_load<KC>(inv_s, KP, w_k[0], K, lane, a_row);
                    const float di = reg_cholesky<KC>(a_row, lane, vec_s);           // vec | dinv: >= 64 floats (the gradient's vec_s is dead by now)
                    grad_k[0] = reg_chol_solve<KC>(a_row, di, grad_k[0], lane, L_s);   // L_s: >= 32 * 33 floats (ctm_l_floats)
                } else {
                    warp_cholesky<R>(inv_s, L_s, dinv_s, w_k, K, KP, lane);
                    warp_chol_solve<R>(L_s, dinv_s, grad_k, K, KP, lane);
                }
#pragma unroll
                for (int r = 0; r < R; r++)
                    if (lane + 32 * r < K) lam_k[r] += grad_k[r];
                if (sqrtf(gn) < p.ntol) break;
            }

            // ---- update_vsq! (fCTM.jl:200-219)
#pragma unroll
            for (int r = 0; r < R; r++) {
                bool active = lane + 32 * r < K;
                for (int it = 0; it < p.niter; it++) {
                    if (active) {
                        float rho = 1.0f;
                        const float ex = Cd * __expf(fmaf(0.5f, vsq_k[r], lam_k[r]) - logzeta);
                        const float grad = -0.5f * (isd_k[r] + ex - 1.0f / vsq_k[r]);
                        const float invhess = -1.0f / (0.25f * ex + 0.5f / (vsq_k[r] * vsq_k[r]));
                        const float pp = invhess * grad;
                        while (vsq_k[r] - rho * pp <= 0.0f) rho *= 0.5f;
                        vsq_k[r] -= rho * pp;
                        if (rho * fabsf(grad) < p.ntol) active = false;
                    }
                    if (!__any_sync(0xffffffffu, active)) break;
                }
                vsq_k[r] += TMVB_EPS;
            }
            float dl = 0.0f;
#pragma unroll
            for (int r = 0; r < R; r++)
                if (lane + 32 * r < K) dl = fmaf(lam_k[r] - lold_k[r], lam_k[r] - lold_k[r], dl);
            dl = warp_sum(dl);
            v++;
            if (sqrtf(dl) < p.vtol || v >= p.viter) break;  // fCTM.jl:264
        }

        // update_beta!(model, d), update_kappa!(model, d): the last phi from (tau_old, e_s = the lambda it was computed from)
        if (!(p.dbg & 1)) {
            f32x2 E01[CPL], E23[CPL];
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                const bool in = kl + LPT * m < CH;
                const float4 E = in ? reinterpret_cast<const float4 *>(e_s)[kl + LPT * m] : make_float4(kPadE, kPadE, kPadE, kPadE);
                E01[m] = pk2(E.x, E.y);
                E23[m] = pk2(E.z, E.w);
            }
            for (int r = 0; r < rounds; r++) {
                const int n = r * S + ts;
                const bool ok = n < Nd;
                const int nn = ok ? n : 0;
                const bool in_tile = nn < cap;
                const int term = in_tile ? term_s[nn] : __ldg(p.terms + o + nn);
                ulonglong2 b[CPL];
                flda_load_row<LPT, CPL>(tile, p.L, RS, K_ld, CH, nn, cap, term, kl, b);
                const float c = in_tile ? cnt_s[nn] : __ldg(p.counts + o + nn);
                const float tauo = __ldcg(p.tau_old + o + nn), tauf = __ldcg(p.tau + o + nn);
                f32x2 p01[CPL], p23[CPL];
                float sn, q;
                flda_row<CPL>(b, E01, E23, tauo, p01, p23, sn, q);
                sn = group_sum<LPT>(sn);
                if (ok) {
                    const float w = tauf * c * rcp_ftz(sn);
                    const f32x2 w2 = pk2(w, w);
                    float *srow = p.stats + (size_t)term * K_ld + 4 * kl;
#pragma unroll
                    for (int m = 0; m < CPL; m++)
                        if (4 * (kl + LPT * m) < K) {
                            float px, py, pz, pw;
                            unpk2(mul2(p01[m], w2), px, py);
                            unpk2(mul2(p23[m], w2), pz, pw);
                            red_add_v4(srow + 4 * LPT * m, px, py, pz, pw);
                        }
                    if (kl == 0) red_add(p.kstats + term, (1.0f - tauf) * c);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            if (i < K_ld) {
                const bool ok = i < K;
                p.lambda[(size_t)d * K_ld + i] = ok ? lam_k[r] : 0.0f;
                p.lambda_old[(size_t)d * K_ld + i] = ok ? lold_k[r] : 0.0f;
                p.vsq[(size_t)d * K_ld + i] = ok ? vsq_k[r] : 0.0f;
            }
        }
        if (lane == 0) {
            p.logzeta[d] = logzeta;
            sweeps_thr += (unsigned long long)v;
        }
    }
    if (lane == 0 && sweeps_thr) atomicAdd(p.small + 1, (double)sweeps_thr);
}

// update_elbo! (fCTM.jl:67-130), literally, one warp per document, lane = topic: the lagged phi from (tau_old, beta_old, lambda_old)
// (fCTM.jl:125), everything else current.  L_old / p.L are the log2 tables of beta_old / beta.
__global__ void fctm_elbo_kernel(const CtmDev p, const float *__restrict__ L_old, const float *__restrict__ kappa, double ln_eta, double ln_1m_eta,
                                 double *out)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int K = p.K, KP = p.KP;
    double acc = 0.0;
    for (long long d = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); d < p.M; d += (long long)gridDim.x * wpb) {
        const long long o = p.doc_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o);
        const float *lam = p.lambda + d * p.K_ld, *lo = p.lambda_old + d * p.K_ld, *vs = p.vsq + d * p.K_ld;
        const float lz = p.logzeta[d];
        float lmax = -INFINITY;
        for (int i = lane; i < K; i += 32) lmax = fmaxf(lmax, lo[i]);
        lmax = warp_max(lmax);
        double dacc = 0.0;
        float Cd = 0.0f, tc = 0.0f;
        for (int n = 0; n < Nd; n++) {
            const int term = p.terms[o + n];
            const float c = p.counts[o + n], tau = p.tau[o + n], tauo = p.tau_old[o + n];
            Cd += c;
            tc = fmaf(tau, c, tc);
            const float *bo = L_old + (size_t)term * p.K_ld, *bn = p.L + (size_t)term * p.K_ld;
            float s = 0.0f;
            for (int i = lane; i < K; i += 32) s += ex2_ftz(fmaf(tauo, bo[i], (lo[i] - lmax) * kLog2e));
            s = warp_sum(s);
            const float l2s = lg2_ftz(s);
            float a = 0.0f;
            for (int i = lane; i < K; i += 32) {
                const float x = fmaf(tauo, bo[i], (lo[i] - lmax) * kLog2e);
                const float ph = ex2_ftz(x) / s;
                // phi (lambda_i + tau ln(beta_i + eps) - ln phi),  ln phi = ln 2 (x - log2 s)
                if (ph > 0.0f) a += ph * (lam[i] + kLn2 * (tau * bn[i] - (x - l2s)));
            }
            a = warp_sum(a);
            a = fmaf(1.0f - tau, logf(kappa[term] + TMVB_EPS), a);          // corpus part of Elogpw, fCTM.jl:90
            const float t0 = 1.0f - tau;
            if (t0 != 0.0f && t0 != 1.0f) a -= t0 * logf(t0) + tau * logf(tau);   // -Elogqc, fCTM.jl:101-105
            if (lane == 0) dacc += (double)(c * a);
        }
        float q = 0.0f, se = 0.0f, lv = 0.0f, dv = 0.0f;
        for (int i = lane; i < K; i += 32) {
            float t = 0.0f;
            for (int j = 0; j < K; j++) t += p.invsigma[i * KP + j] * (lam[j] - p.mu[j]);
            q += (lam[i] - p.mu[i]) * t;
            dv += p.invsigma[i * KP + i] * vs[i];
            se += expf(lam[i] + 0.5f * vs[i] - lz);
            lv += logf(vs[i]);
        }
        dacc += (double)(-0.5f * (dv + q) + 0.5f * lv);
        dacc = warp_sum_d(dacc);
        se = warp_sum(se);
        if (lane == 0) {
            const double elogpc = log(exp((double)tc * ln_eta + ((double)Cd - (double)tc) * ln_1m_eta) + TMVB_EPS_D);   // fCTM.jl:73-77
            acc += dacc - (double)Cd * ((double)se + (double)lz - 1.0) + elogpc;
        }
    }
    if (lane == 0 && acc != 0.0) atomicAdd(out, acc);
}

// Second moments for update_sigma!/update_mu! (CTM.jl:102-111): mom = [sum lambda (K_ld) | sum vsq (K_ld) | sum lambda lambda' (K x K)]
template <int NQ>   // pairs per thread: K*K <= 256 NQ
__global__ void ctm_moments_kernel(const float *__restrict__ lambda, const float *__restrict__ vsq, long long M, int K, int K_ld, double *__restrict__ mom)
{
    extern __shared__ float lam_s[];  // [chunk][K_ld]
    const int CHUNK = 32;
    const int npairs = K * K;
    double acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) acc[q] = 0.0;
    double sl = 0.0, sv = 0.0;
    for (long long d0 = (long long)blockIdx.x * CHUNK; d0 < M; d0 += (long long)gridDim.x * CHUNK) {
        const int nd = (int)min((long long)CHUNK, M - d0);
        __syncthreads();
        for (int q = threadIdx.x; q < nd * K_ld; q += blockDim.x) lam_s[q] = lambda[d0 * K_ld + q];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const int pr = threadIdx.x + q * 256;
            if (pr >= npairs) break;
            const int i = pr / K, j = pr - i * K;
            double a = 0.0;
            for (int dd = 0; dd < nd; dd++) a += (double)(lam_s[dd * K_ld + i] * lam_s[dd * K_ld + j]);
            acc[q] += a;
        }
        if (threadIdx.x < K)
            for (int dd = 0; dd < nd; dd++) {
                sl += (double)lam_s[dd * K_ld + threadIdx.x];
                sv += (double)vsq[(d0 + dd) * K_ld + threadIdx.x];
            }
    }
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int pr = threadIdx.x + q * 256;
        if (pr < npairs && acc[q] != 0.0) atomicAdd(mom + 2 * K_ld + pr, acc[q]);
    }
    if (threadIdx.x < K) {
        atomicAdd(mom + threadIdx.x, sl);
        atomicAdd(mom + K_ld + threadIdx.x, sv);
    }
}

// update_elbo! restated (CTM.jl:56-98) on the device state: phi from beta_old / lambda_old, everything else current.
// One warp per document, lanes over topics; fp32 per element, fp64 accumulation.  logdet(invsigma) enters on the host.
// RM = topics per lane (i = lane + 32 r): exp(lambda_old - max) and lambda are per-document and stay in registers; per (token,
// topic) one product for s, then phi (lambda_i + ln(beta_i + eps) - ln phi_i) with two logarithms.
template <int RM>
__global__ void ctm_elbo_kernel(const CtmDev p, const float *__restrict__ beta_old, double *out)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int K = p.K, KP = p.KP;
    double acc = 0.0;
    for (long long d = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); d < p.M; d += (long long)gridDim.x * wpb) {
        const long long o = p.doc_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o);
        const float *lam = p.lambda + d * p.K_ld, *lo = p.lambda_old + d * p.K_ld, *vs = p.vsq + d * p.K_ld;
        const float lz = p.logzeta[d];
        float e_r[RM], lam_r[RM];
        float lmax = -INFINITY;
#pragma unroll
        for (int r = 0; r < RM; r++) {
            const int i = lane + 32 * r;
            e_r[r] = (i < K) ? lo[i] : -INFINITY;
            lam_r[r] = (i < K) ? lam[i] : 0.0f;
            lmax = fmaxf(lmax, e_r[r]);
        }
        lmax = warp_max(lmax);
#pragma unroll
        for (int r = 0; r < RM; r++) e_r[r] = expf(e_r[r] - lmax);
        double dacc = 0.0;
        float Cd = 0.0f;
        for (int n = 0; n < Nd; n++) {
            const int term = p.terms[o + n];
            const float c = p.counts[o + n];
            Cd += c;
            const float *bo = beta_old + (size_t)term * p.K_ld, *bn = p.beta + (size_t)term * p.K_ld;
            float u[RM], s = 0.0f;
#pragma unroll
            for (int r = 0; r < RM; r++) {
                const int i = lane + 32 * r;
                u[r] = (i < K) ? bo[i] * e_r[r] : 0.0f;
                s += u[r];
            }
            s = warp_sum(s);
            const float rs = 1.0f / s;
            float a = 0.0f;
#pragma unroll
            for (int r = 0; r < RM; r++) {
                const int i = lane + 32 * r;
                const float ph = u[r] * rs;
                if (i < K && ph > 0.0f) a += ph * (lam_r[r] + logf(bn[i] + TMVB_EPS) - logf(ph));
            }
            dacc += (double)(c * a);
        }
        float q = 0.0f, se = 0.0f, lv = 0.0f, dv = 0.0f;
        for (int i = lane; i < K; i += 32) {
            float t = 0.0f;
            for (int j = 0; j < K; j++) t += p.invsigma[i * KP + j] * (lam[j] - p.mu[j]);
            q += (lam[i] - p.mu[i]) * t;
            dv += p.invsigma[i * KP + i] * vs[i];
            se += expf(lam[i] + 0.5f * vs[i] - lz);
            lv += logf(vs[i]);
        }
        dacc += (double)(-0.5f * (dv + q) + 0.5f * lv);
        dacc = warp_sum_d(dacc);
        se = warp_sum(se);
        if (lane == 0) acc += dacc - (double)Cd * ((double)se + (double)lz - 1.0);
    }
    if (lane == 0 && acc != 0.0) atomicAdd(out, acc);
}

// phi[K x sumN] in the caller's token order from beta_old / lambda_old (CTM.jl:93)
__global__ void ctm_phi_kernel(const CtmDev p, const float *__restrict__ beta_old, const long long *__restrict__ src_off, float *__restrict__ phi)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long d = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); d < p.M; d += (long long)gridDim.x * wpb) {
        const long long o = p.doc_off[d], so = src_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o);
        const float *lo = p.lambda_old + d * p.K_ld;
        float lmax = -INFINITY;
        for (int i = lane; i < p.K; i += 32) lmax = fmaxf(lmax, lo[i]);
        lmax = warp_max(lmax);
        for (int n = 0; n < Nd; n++) {
            const float *bo = beta_old + (size_t)p.terms[o + n] * p.K_ld;
            float s = 0.0f;
            for (int i = lane; i < p.K; i += 32) s += bo[i] * expf(lo[i] - lmax);
            s = warp_sum(s);
            for (int i = lane; i < p.K; i += 32) phi[(size_t)(so + n) * p.K + i] = bo[i] * expf(lo[i] - lmax) / s;
        }
    }
}

typedef void (*CtmEstepFn)(const CtmDev, int, int, int, int, int *);
// layouts with R = ceil(LPT*CPL/8) <= 4, i.e. K <= 128; the others are not instantiated
template <int L, int C, bool E, bool OK = ((L * C + 7) / 8 <= 4)>
struct CtmPick {
    static CtmEstepFn get() { return (CtmEstepFn)ctm_estep_kernel<L, C, E>; }
};
template <int L, int C, bool E>
struct CtmPick<L, C, E, false> {
    static CtmEstepFn get() { return nullptr; }
};
#define TMVB_CTM_FN(L, C) {CtmPick<L, C, false>::get(), CtmPick<L, C, true>::get()},
static CtmEstepFn ctm_fn(int layout, int elbo)
{
    static const CtmEstepFn tab[kNumLaneLayouts][2] = {TMVB_FOR_EACH_LAYOUT(TMVB_CTM_FN)};
    return tab[layout][elbo];
}

template <int L, int C, bool OK = ((L * C + 7) / 8 <= 4)>
struct FctmPick {
    static CtmEstepFn get() { return (CtmEstepFn)fctm_estep_kernel<L, C>; }
};
template <int L, int C>
struct FctmPick<L, C, false> {
    static CtmEstepFn get() { return nullptr; }
};
#define TMVB_FCTM_FN(L, C) FctmPick<L, C>::get(),
static CtmEstepFn fctm_fn(int layout)
{
    static const CtmEstepFn tab[kNumLaneLayouts] = {TMVB_FOR_EACH_LAYOUT(TMVB_FCTM_FN)};
    return tab[layout];
}

// in-place inverse and log-determinant of an SPD matrix (fp64, host): inv(sigma), logdet (CTM.jl:57,110)
static int spd_inv_logdet(int K, std::vector<double> &A, double *logdet)
{
    std::vector<double> L(A);
    for (int j = 0; j < K; j++) {
        double s = L[j * K + j];
        for (int k = 0; k < j; k++) s -= L[j * K + k] * L[j * K + k];
        if (!(s > 0.0)) return -1;
        const double l = sqrt(s);
        L[j * K + j] = l;
        for (int i = j + 1; i < K; i++) {
            double t = L[i * K + j];
            for (int k = 0; k < j; k++) t -= L[i * K + k] * L[j * K + k];
            L[i * K + j] = t / l;
        }
    }
    double ld = 0.0;
    for (int i = 0; i < K; i++) ld += 2.0 * log(L[i * K + i]);
    if (logdet) *logdet = ld;
    std::vector<double> e(K);
    for (int c = 0; c < K; c++) {
        for (int i = 0; i < K; i++) e[i] = (i == c) ? 1.0 : 0.0;
        for (int i = 0; i < K; i++) {
            double t = e[i];
            for (int k = 0; k < i; k++) t -= L[i * K + k] * e[k];
            e[i] = t / L[i * K + i];
        }
        for (int i = K - 1; i >= 0; i--) {
            double t = e[i];
            for (int k = i + 1; k < K; k++) t -= L[k * K + i] * e[k];
            e[i] = t / L[i * K + i];
        }
        for (int i = 0; i < K; i++) A[i * K + c] = e[i];
    }
    return 0;
}

}  // namespace tmvb

using namespace tmvb;

struct tmvb_ctm_s {
    Shard s;
    int KP = 0;
    bool no_scatter = false;   // predict: the E-step leaves the statistics alone
    bool elbo_valid = false;
    float *d_lambda = nullptr, *d_lambda_old = nullptr, *d_vsq = nullptr, *d_logzeta = nullptr;
    float *d_mu = nullptr, *d_invsigma = nullptr;
    std::vector<double> h_mu, h_sigma, h_invsigma;  // fp64 masters (K, K x K, K x K)
    double h_logdet_inv = 0.0;
    double *d_small = nullptr;  // [0] ELBO docs | [1] sweeps | [2..] moments: sum lambda (K_ld) | sum vsq (K_ld) | sum lambda lambda' (K*K)
    double *d_local = nullptr;  // [K_ld] rowsum | [K_ld] elbo_w | [1] standalone ELBO
    std::vector<double> h_mom;  // host copy of the (reduced) moments of the last E-step
    int64_t n_small = 0;
    // filtered CTM (tmvb_fctm_*): fCTM.jl's eta / kappa / tau on top of the CTM state
    bool filtered = false, filt_set = false;
    double eta = 0.5;
    float *d_L[2] = {nullptr, nullptr}, *d_kappa = nullptr, *d_kappa_old = nullptr, *d_kstats = nullptr, *d_kq = nullptr;
    float *d_tau = nullptr, *d_tau_old = nullptr;
    size_t tau_cap = 0;
    Comm comm;   // peer-memory all-reduce of the statistics (multi-GPU)
};

namespace {

CtmDev ctm_view(tmvb_ctm_t h)
{
    Shard &s = h->s;
    CtmDev p;
    memset(&p, 0, sizeof(p));  // the struct is also the key of the captured launch graph: no indeterminate padding
    p.K = (int)s.K;
    p.K_ld = s.K_ld;
    p.V = (int)s.V;
    p.RS = s.RS;
    p.KP = h->KP;
    p.M = s.M;
    p.beta = s.d_beta[s.cur];
    p.stats = s.d_stats;
    p.doc_off = s.d_doc_off;
    p.terms = s.d_terms;
    p.counts = s.d_counts;
    p.lambda = h->d_lambda;
    p.lambda_old = h->d_lambda_old;
    p.vsq = h->d_vsq;
    p.logzeta = h->d_logzeta;
    p.mu = h->d_mu;
    p.invsigma = h->d_invsigma;
    p.small = h->d_small;
    p.L = h->d_L[s.cur];
    p.kq = h->d_kq;
    p.kstats = h->d_kstats;
    p.tau = h->d_tau;
    p.tau_old = h->d_tau_old;
    p.niter = 0;
    p.ntol = 0.f;
    p.viter = 0;
    p.vtol = 0.f;
    p.stage_bulk = env_int("TMVB_STAGE_BULK", 1);
    p.dbg = env_int("TMVB_DBG", 0) | (h->no_scatter ? 1 : 0);   // bit 0: the scatter pass computes but does not store (predict)
    return p;
}

void ctm_free(tmvb_ctm_t h)
{
    cudaSetDevice(h->s.device);
    if (h->s.stream) cudaStreamSynchronize(h->s.stream);
    cudaFree(h->d_lambda);
    cudaFree(h->d_lambda_old);
    cudaFree(h->d_vsq);
    cudaFree(h->d_logzeta);
    cudaFree(h->d_mu);
    cudaFree(h->d_invsigma);
    cudaFree(h->d_small);
    cudaFree(h->d_local);
    cudaFree(h->d_L[0]);
    cudaFree(h->d_L[1]);
    cudaFree(h->d_kappa);
    cudaFree(h->d_kappa_old);
    cudaFree(h->d_kstats);
    cudaFree(h->d_kq);
    cudaFree(h->d_tau);
    cudaFree(h->d_tau_old);
    comm_free(&h->comm);
    shard_free(&h->s);
}

// push the fp64 masters mu / invsigma to the device (`@buffer model.invsigma`, macros.jl:67)
PeerReduce ctm_peer_bufs(tmvb_ctm_t h)
{
    PeerReduce b;
    b.f[0] = h->s.d_stats;
    b.nf[0] = (long long)h->s.V * h->s.K_ld;
    if (h->filtered) {
        b.f[1] = h->d_kstats;
        b.nf[1] = ((long long)std::max<int64_t>(h->s.V, 1) + 3) / 4 * 4;
    }
    b.small = h->d_small;
    b.n_small = h->n_small;
    return b;
}

int ctm_push_globals(tmvb_ctm_t h)
{
    Shard &s = h->s;
    const int K = (int)s.K, KP = h->KP;
    std::vector<float> inv((size_t)K * KP, 0.f), mu(s.K_ld, 0.f);
    for (int i = 0; i < K; i++) {
        mu[i] = (float)h->h_mu[i];
        for (int j = 0; j < K; j++) inv[(size_t)i * KP + j] = (float)h->h_invsigma[(size_t)i * K + j];
    }
    TMVB_CUDA(cudaMemcpyAsync(h->d_invsigma, inv.data(), inv.size() * 4, cudaMemcpyHostToDevice, s.stream));
    TMVB_CUDA(cudaMemcpyAsync(h->d_mu, mu.data(), mu.size() * 4, cudaMemcpyHostToDevice, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    s.st.h2d_bytes += (int64_t)K * K * 4 + K * 4;
    return 0;
}

}  // namespace

extern "C" {

int tmvb_ctm_create(tmvb_ctm_t *out, int64_t K, int64_t M, int64_t V, int device, void *stream)
{
    TMVB_CHECK_ARG(out != nullptr, "handle pointer is NULL");
    *out = nullptr;
    TMVB_CHECK_ARG(K > 0, "number of topics must be a positive integer");  // gpuCTM.jl:55
    if (K > 128) return fail(-2, "gpuCTM supports K <= 128 (K=%lld): the K x K Cholesky factor of update_lambda! lives in shared memory", (long long)K);
    tmvb_ctm_t h = new tmvb_ctm_s();
    const int64_t K_ld = (K + 7) / 8 * 8;
    h->n_small = 2 + 2 * K_ld + K * K;
    int rc = shard_create(&h->s, K, M, V, device, stream, (size_t)h->n_small + 2 * K_ld + 8);
    if (rc == 0) {
        Shard &s = h->s;
        h->KP = ctm_kp((int)K);
        const size_t km = (size_t)std::max<int64_t>(M, 1) * s.K_ld;
        cudaError_t e = cudaSuccess;
        auto A = [&](void **p, size_t bytes) {
            if (e == cudaSuccess) e = cudaMalloc(p, bytes);
            if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes, s.stream);
        };
        A((void **)&h->d_lambda, km * 4);
        A((void **)&h->d_lambda_old, km * 4);
        A((void **)&h->d_vsq, km * 4);
        A((void **)&h->d_logzeta, (size_t)std::max<int64_t>(M, 1) * 4);
        A((void **)&h->d_mu, s.K_ld * 4);
        A((void **)&h->d_invsigma, (size_t)K * h->KP * 4);
        A((void **)&h->d_small, h->n_small * 8);
        A((void **)&h->d_local, (2 * s.K_ld + 2) * 8);
        for (int eb = 0; eb < 2 && e == cudaSuccess; eb++) {
            CtmEstepFn fn = ctm_fn(s.layout, eb);
            if (!fn) {
                rc = fail(-2, "internal: no CTM kernel for this K");
                break;
            }
            e = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem_optin);
        }
        if (e != cudaSuccess) rc = fail((int)e, "device allocation failed: %s", cudaGetErrorString(e));
    }
    if (rc != 0) {
        ctm_free(h);
        delete h;
        return rc;
    }
    // gpuCTM.jl:63-65: mu = 0, sigma = invsigma = I
    h->h_mu.assign(K, 0.0);
    h->h_sigma.assign((size_t)K * K, 0.0);
    for (int i = 0; i < K; i++) h->h_sigma[(size_t)i * K + i] = 1.0;
    h->h_invsigma = h->h_sigma;
    h->h_logdet_inv = 0.0;
    h->h_mom.assign(h->n_small, 0.0);
    *out = h;
    return 0;
}

int tmvb_ctm_destroy(tmvb_ctm_t h)
{
    if (!h) return 0;
    ctm_free(h);
    delete h;
    return 0;
}

int tmvb_ctm_set_corpus(tmvb_ctm_t h, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    // a small tile: the E-step is bound by the latency of the per-document Newton / Cholesky chains, not by the token passes, so
    // resident documents per SM beat shared-memory reads of the rows (CiteULike K=30: 6.5 ms with full tiles, 4.0 ms with 16 tokens)
    h->s.tile_cap_max = 16;
    return shard_set_corpus(&h->s, N_cumsum, terms, counts, ctm_fixed_smem(h->s.RS, h->s.lpt, (int)h->s.K, h->s.K_ld));
}

int tmvb_ctm_upload(tmvb_ctm_t h, const float *mu, const float *sigma, const float *beta, const float *lambda, const float *vsq,
                    const float *logzeta)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    const int K = (int)s.K;
    if (mu)
        for (int i = 0; i < K; i++) {
            if (!isfinite(mu[i])) return fail(-5, "mu must be finite.");  // modelutils.jl:287
            h->h_mu[i] = (double)mu[i];
        }
    if (sigma) {
        for (int q = 0; q < K * K; q++) h->h_sigma[q] = (double)sigma[q];
        h->h_invsigma = h->h_sigma;
        double ld = 0.0;
        if (spd_inv_logdet(K, h->h_invsigma, &ld)) return fail(-5, "sigma must be positive-definite.");  // modelutils.jl:289
        h->h_logdet_inv = -ld;
    }
    TMVB_TRY(ctm_push_globals(h));
    if (beta && s.V > 0) {
        TMVB_TRY(shard_upload_rows(&s, beta, s.d_beta[s.cur], s.V, nullptr, 0));
        TMVB_TRY(shard_check_stochastic(&s, s.d_beta[s.cur]));   // the row sums of check_model, on the device copy
        TMVB_CUDA(cudaMemcpyAsync(s.d_beta[s.cur ^ 1], s.d_beta[s.cur], (size_t)s.V * s.K_ld * 4, cudaMemcpyDeviceToDevice, s.stream));  // CTM.jl:42
        if (h->filtered) {
            TMVB_TRY(filt_log_table(&s, s.d_beta[0], h->d_L[0]));
            TMVB_TRY(filt_log_table(&s, s.d_beta[1], h->d_L[1]));
        }
    }
    if ((lambda || vsq || logzeta) && s.M > 0) {
        TMVB_CHECK_ARG(s.corpus_set, "set_corpus must precede the upload of per-document parameters");
        TMVB_TRY(shard_upload_rows(&s, lambda, h->d_lambda, s.M, s.d_perm, 3));
        if (lambda)  // lambda_old = deepcopy(lambda), CTM.jl:45
            TMVB_CUDA(cudaMemcpyAsync(h->d_lambda_old, h->d_lambda, (size_t)s.M * s.K_ld * 4, cudaMemcpyDeviceToDevice, s.stream));
        TMVB_TRY(shard_upload_rows(&s, vsq, h->d_vsq, s.M, s.d_perm, 2));
        if (logzeta) {
            std::vector<float> lz(s.M);
            for (int64_t p = 0; p < s.M; p++) {
                lz[p] = logzeta[s.h_perm[p]];
                if (!isfinite(lz[p])) return fail(-5, "logzeta must be finite.");  // modelutils.jl:301
            }
            TMVB_CUDA(cudaMemcpyAsync(h->d_logzeta, lz.data(), s.M * 4, cudaMemcpyHostToDevice, s.stream));
            TMVB_CUDA(cudaStreamSynchronize(s.stream));
            s.st.h2d_bytes += s.M * 4;
        }
    }
    int verr = 0;
    TMVB_TRY(shard_validation(&s, &verr));
    if (verr & 0x4003) return fail(-5, "beta must be a right stochastic matrix.");  // modelutils.jl:293
    if (verr & 0x40) return fail(-5, "lambda must be finite.");                 // modelutils.jl:296
    if (verr & 0x10) return fail(-5, "vsq must be finite.");                    // modelutils.jl:299
    if (verr & 0x20) return fail(-5, "vsq must be positive.");                  // modelutils.jl:300
    h->elbo_valid = false;
    return 0;
}

int tmvb_ctm_estep(tmvb_ctm_t h, int niter, float ntol, int viter, float vtol, int want_elbo)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CHECK_ARG(viter >= 1 && niter >= 0, "iteration parameters must be nonnegative (viter >= 1)");
    TMVB_CHECK_ARG(vtol >= 0.f && ntol >= 0.f, "tolerance parameters must be nonnegative");  // gpuCTM.jl:489
    Shard &s = h->s;
    TMVB_CHECK_ARG(s.corpus_set, "set_corpus has not been called");
    TMVB_CUDA(cudaSetDevice(s.device));
    CtmDev p = ctm_view(h);
    p.niter = niter;
    p.ntol = ntol;
    p.viter = viter;
    p.vtol = vtol;
    TMVB_CUDA(cudaEventRecord(s.ev[0], s.stream));
    TMVB_CUDA(cudaMemsetAsync(h->d_small, 0, h->n_small * 8, s.stream));
    if (h->filtered) {
        TMVB_CHECK_ARG(h->filt_set, "tmvb_fctm_upload has not been called");
        want_elbo = 0;   // the filtered ELBO is evaluated by fctm_elbo_kernel (tmvb_ctm_elbo), not from fused partials
    }
    const void *fn = h->filtered ? (const void *)fctm_fn(s.layout) : (const void *)ctm_fn(s.layout, want_elbo != 0);
    const void *fns[2] = {fn, fn};
    TMVB_TRY(shard_launch(&s, pick_by_warps, fns, &p, sizeof(p)));
    if (s.M > 0) {
        const int grid = (int)std::min<int64_t>((s.M + 31) / 32, (int64_t)s.n_sm * 4);
        if (s.K <= 64)
            ctm_moments_kernel<16><<<grid, 256, 32 * s.K_ld * 4, s.stream>>>(h->d_lambda, h->d_vsq, s.M, (int)s.K, s.K_ld, h->d_small + 2);
        else
            ctm_moments_kernel<64><<<grid, 256, 32 * s.K_ld * 4, s.stream>>>(h->d_lambda, h->d_vsq, s.M, (int)s.K, s.K_ld, h->d_small + 2);
        TMVB_CUDA(cudaGetLastError());
        s.st.kernel_launches++;
    }
    TMVB_CUDA(cudaEventRecord(s.ev[1], s.stream));
    s.estep_timed = true;
    h->elbo_valid = (want_elbo != 0);
    return 0;
}

// the inner loop of predict (modelutils.jl:903-913): the E-step without update_beta!(model, d) -- no statistics are scattered
int tmvb_ctm_predict(tmvb_ctm_t h, int niter, float ntol, int viter, float vtol)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    h->no_scatter = true;
    const int rc = tmvb_ctm_estep(h, niter, ntol, viter, vtol, 0);
    h->no_scatter = false;
    return rc;
}

int tmvb_ctm_reduce_buffers(tmvb_ctm_t h, void **stats, int64_t *n_stats, void **small, int64_t *n_small)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    if (stats) *stats = h->s.d_stats;
    if (n_stats) *n_stats = (int64_t)h->s.V * h->s.K_ld;
    if (small) *small = h->d_small;
    if (n_small) *n_small = h->n_small;
    return 0;
}

int tmvb_ctm_mstep(tmvb_ctm_t h, int64_t M_total)
{
    TMVB_CHECK_ARG(h != nullptr && M_total > 0, "bad arguments");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    TMVB_CUDA(cudaEventRecord(s.ev[2], s.stream));
    TMVB_TRY(shard_normalize(&s, h->d_local, h->elbo_valid, false));  // update_beta! CTM.jl:114-118
    if (h->filtered) {   // the log2 table of the new beta; update_kappa!(model) (fCTM.jl:154-158)
        TMVB_TRY(filt_log_table(&s, s.d_beta[s.cur], h->d_L[s.cur]));
        TMVB_TRY(filt_kappa_update(&s, h->d_kstats, h->d_kappa, h->d_kappa_old));
        TMVB_TRY(filt_push_kq(&s, h->d_kappa, h->d_kq, h->eta));   // eta is not updated: update_eta! is commented out, fCTM.jl:279
    }
    TMVB_CUDA(cudaMemcpyAsync(s.h_pinned, h->d_small, h->n_small * 8, cudaMemcpyDeviceToHost, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    s.st.d2h_bytes += h->n_small * 8;
    memcpy(h->h_mom.data(), s.h_pinned, h->n_small * 8);
    if (h->comm.connected) {   // a peer that never arrived: raise instead of normalising half-summed statistics
        int pst = 0;
        TMVB_TRY(peer_status(&h->comm, s.stream, &pst));
    }
    const int K = (int)s.K, K_ld = s.K_ld;
    const double *sl = h->h_mom.data() + 2, *sv = sl + K_ld, *G = sv + K_ld;
    const double Md = (double)M_total;
    // update_sigma! (CTM.jl:108-111) with the OLD mu: sum (l-mu)(l-mu)' = G - sl mu' - mu sl' + M mu mu'
    for (int i = 0; i < K; i++)
        for (int j = 0; j < K; j++) {
            double v = G[i * K + j] - sl[i] * h->h_mu[j] - h->h_mu[i] * sl[j] + Md * h->h_mu[i] * h->h_mu[j];
            if (i == j) v += sv[i];
            h->h_sigma[(size_t)i * K + j] = v / Md;
        }
    for (int i = 0; i < K; i++)  // symmetrise against fp noise in G
        for (int j = 0; j < i; j++) h->h_sigma[(size_t)i * K + j] = h->h_sigma[(size_t)j * K + i] = 0.5 * (h->h_sigma[(size_t)i * K + j] + h->h_sigma[(size_t)j * K + i]);
    h->h_invsigma = h->h_sigma;
    double ld = 0.0;
    if (spd_inv_logdet(K, h->h_invsigma, &ld)) return fail(-5, "sigma must be positive-definite.");
    h->h_logdet_inv = -ld;
    for (int i = 0; i < K; i++) h->h_mu[i] = sl[i] / Md;  // update_mu! CTM.jl:102-104
    TMVB_TRY(ctm_push_globals(h));
    TMVB_CUDA(cudaEventRecord(s.ev[3], s.stream));
    s.mstep_timed = true;
    return 0;
}

int tmvb_ctm_elbo(tmvb_ctm_t h, int mode, int64_t M_total, double *elbo_docs, double *elbo_global)
{
    TMVB_CHECK_ARG(h && elbo_docs && elbo_global, "NULL argument");
    TMVB_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 or 1");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    const int K = (int)s.K, K_ld = s.K_ld;
    const double Md = (double)M_total;
    const double LOG2PI = 1.8378770664093453;
    if (h->filtered) mode = 1;
    if (mode == 0) {
        TMVB_CHECK_ARG(h->elbo_valid, "mode 0 needs estep(want_elbo=1) followed by mstep");
        TMVB_CUDA(cudaMemcpyAsync(s.h_pinned, h->d_local + K_ld, 8, cudaMemcpyDeviceToHost, s.stream));
        TMVB_CUDA(cudaStreamSynchronize(s.stream));
        s.st.d2h_bytes += 8;
        const double *sl = h->h_mom.data() + 2, *sv = sl + K_ld, *G = sv + K_ld;
        // sum_d Elogpeta (CTM.jl:56-59) with the new mu / invsigma, from the moments
        double tr = 0.0, dv = 0.0;
        for (int i = 0; i < K; i++) {
            dv += h->h_invsigma[(size_t)i * K + i] * sv[i];
            for (int j = 0; j < K; j++) {
                const double S = G[i * K + j] - sl[i] * h->h_mu[j] - h->h_mu[i] * sl[j] + Md * h->h_mu[i] * h->h_mu[j];
                tr += h->h_invsigma[(size_t)i * K + j] * S;
            }
        }
        double g = 0.5 * (Md * h->h_logdet_inv - Md * K * LOG2PI - dv - tr);
        g += 0.5 * Md * K * (LOG2PI + 1.0);  // constant part of entropy(MvNormal), CTM.jl:77
        g += s.h_pinned[0];                  // Elogpw over the statistics
        *elbo_docs = h->h_mom[0];
        *elbo_global = g;
        return 0;
    }
    TMVB_CHECK_ARG(s.corpus_set, "set_corpus has not been called");
    CtmDev p = ctm_view(h);
    double *out = h->d_local + 2 * K_ld;
    TMVB_CUDA(cudaMemsetAsync(out, 0, 8, s.stream));
    if (s.M > 0) {
        if (h->filtered)
            fctm_elbo_kernel<<<grid_for(s.M * 32, 128, s.n_sm), 128, 0, s.stream>>>(p, h->d_L[s.cur ^ 1], h->d_kappa, log(h->eta), log1p(-h->eta), out);
        else
        {
            const int grid = grid_for(s.M * 32, 128, s.n_sm);
            if (K <= 32) ctm_elbo_kernel<1><<<grid, 128, 0, s.stream>>>(p, s.d_beta[s.cur ^ 1], out);
            else if (K <= 64) ctm_elbo_kernel<2><<<grid, 128, 0, s.stream>>>(p, s.d_beta[s.cur ^ 1], out);
            else ctm_elbo_kernel<4><<<grid, 128, 0, s.stream>>>(p, s.d_beta[s.cur ^ 1], out);
        }
        TMVB_CUDA(cudaGetLastError());
        s.st.kernel_launches++;
    }
    TMVB_CUDA(cudaMemcpyAsync(s.h_pinned, out, 8, cudaMemcpyDeviceToHost, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    s.st.d2h_bytes += 8;
    // per-document constants: 0.5 (logdet invsigma - K log 2pi) + 0.5 K (log 2pi + 1), for this shard's documents
    *elbo_docs = s.h_pinned[0] + (double)s.M * 0.5 * (h->h_logdet_inv - K * LOG2PI + K * (LOG2PI + 1.0));
    *elbo_global = 0.0;
    return 0;
}

int tmvb_ctm_download(tmvb_ctm_t h, float *mu, float *sigma, float *invsigma, float *beta, float *lambda, float *vsq, float *logzeta)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    const int K = (int)s.K;
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    if (mu)
        for (int i = 0; i < K; i++) mu[i] = (float)h->h_mu[i];
    if (sigma)
        for (int q = 0; q < K * K; q++) sigma[q] = (float)h->h_sigma[q];
    if (invsigma)
        for (int q = 0; q < K * K; q++) invsigma[q] = (float)h->h_invsigma[q];
    TMVB_TRY(shard_download_rows(&s, s.d_beta[s.cur], beta, s.V, nullptr));
    TMVB_TRY(shard_download_rows(&s, h->d_lambda, lambda, s.M, s.d_perm));
    TMVB_TRY(shard_download_rows(&s, h->d_vsq, vsq, s.M, s.d_perm));
    if (logzeta && s.M > 0) {
        std::vector<float> lz(s.M);
        TMVB_CUDA(cudaMemcpyAsync(lz.data(), h->d_logzeta, s.M * 4, cudaMemcpyDeviceToHost, s.stream));
        TMVB_CUDA(cudaStreamSynchronize(s.stream));
        for (int64_t p = 0; p < s.M; p++) logzeta[s.h_perm[p]] = lz[p];
        s.st.d2h_bytes += s.M * 4;
    }
    return 0;
}

int tmvb_ctm_download_old(tmvb_ctm_t h, float *beta_old, float *lambda_old)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    TMVB_TRY(shard_download_rows(&s, s.d_beta[s.cur ^ 1], beta_old, s.V, nullptr));
    TMVB_TRY(shard_download_rows(&s, h->d_lambda_old, lambda_old, s.M, s.d_perm));
    return 0;
}

int tmvb_ctm_materialize_phi(tmvb_ctm_t h, float *phi)
{
    TMVB_CHECK_ARG(h && phi, "NULL argument");
    Shard &s = h->s;
    TMVB_CHECK_ARG(s.corpus_set, "set_corpus has not been called");
    TMVB_CUDA(cudaSetDevice(s.device));
    if (s.nnz == 0) return 0;
    const size_t bytes = (size_t)s.nnz * s.K * 4;
    TMVB_TRY(shard_scratch(&s, bytes));
    CtmDev p = ctm_view(h);
    ctm_phi_kernel<<<grid_for(s.M * 32, 128, s.n_sm), 128, 0, s.stream>>>(p, s.d_beta[s.cur ^ 1], s.d_src_off, (float *)s.d_scratch);
    TMVB_CUDA(cudaGetLastError());
    s.st.kernel_launches++;
    TMVB_CUDA(cudaMemcpyAsync(phi, s.d_scratch, bytes, cudaMemcpyDeviceToHost, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    s.st.d2h_bytes += bytes;
    return 0;
}

int tmvb_ctm_topics(tmvb_ctm_t h, int32_t *topics)
{
    TMVB_CHECK_ARG(h && topics, "NULL argument");
    TMVB_CUDA(cudaSetDevice(h->s.device));
    return shard_topics(&h->s, h->s.d_beta[h->s.cur], nullptr, topics);
}

/* ---- filtered CTM (src/fCTM.jl): a CTM handle with eta / kappa / tau on top; every tmvb_ctm_* call applies to it ---- */
int tmvb_fctm_create(tmvb_ctm_t *out, int64_t K, int64_t M, int64_t V, int device, void *stream)
{
    TMVB_TRY(tmvb_ctm_create(out, K, M, V, device, stream));
    tmvb_ctm_t h = *out;
    Shard &s = h->s;
    h->filtered = true;
    const size_t kv = (size_t)std::max<int64_t>(V, 1) * s.K_ld, v1 = (size_t)std::max<int64_t>(V, 1);
    cudaError_t e = cudaSuccess;
    auto A = [&](void **p, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes, s.stream);
    };
    A((void **)&h->d_L[0], kv * 4);
    A((void **)&h->d_L[1], kv * 4);
    A((void **)&h->d_kappa, v1 * 4);
    A((void **)&h->d_kappa_old, v1 * 4);
    A((void **)&h->d_kstats, (v1 + 3) / 4 * 16);   // a multiple of four floats: the peer all-reduce moves 16-byte elements
    A((void **)&h->d_kq, (v1 + 1) * 4);
    if (e == cudaSuccess && !fctm_fn(s.layout)) {
        tmvb_ctm_destroy(h);
        *out = nullptr;
        return fail(-2, "internal: no fCTM kernel for this K");
    }
    if (e == cudaSuccess) e = cudaFuncSetAttribute((const void *)fctm_fn(s.layout), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem_optin);
    if (e != cudaSuccess) {
        tmvb_ctm_destroy(h);
        *out = nullptr;
        return fail((int)e, "device allocation failed: %s", cudaGetErrorString(e));
    }
    return 0;
}

/* eta, kappa[V], tau[sum N] (the caller's CSR order) of fCTM.jl:10-28; *_old are set to the uploaded values (fCTM.jl:50,59).
 * Call after tmvb_ctm_set_corpus; any pointer may be NULL. */
int tmvb_fctm_upload(tmvb_ctm_t h, const double *eta, const float *kappa, const float *tau)
{
    TMVB_CHECK_ARG(h != nullptr && h->filtered, "not a filtered-CTM handle");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    if (eta) {
        if (!(*eta >= 0.0 && *eta <= 1.0)) return fail(-5, "eta must belong to the interval [0,1].");   // modelutils.jl:104
        h->eta = *eta;
    }
    if (kappa && s.V > 0) {
        double ks = 0.0;
        for (int64_t j = 0; j < s.V; j++) {
            if (!(kappa[j] >= 0.f) || !isfinite(kappa[j])) return fail(-5, "kappa must be a probability vector.");
            ks += kappa[j];
        }
        if (fabs(ks - 1.0) > 1e-3) return fail(-5, "kappa must be a probability vector.");
        TMVB_CUDA(cudaMemcpyAsync(h->d_kappa, kappa, s.V * 4, cudaMemcpyHostToDevice, s.stream));
        TMVB_CUDA(cudaMemcpyAsync(h->d_kappa_old, h->d_kappa, s.V * 4, cudaMemcpyDeviceToDevice, s.stream));
        TMVB_CUDA(cudaStreamSynchronize(s.stream));
        s.st.h2d_bytes += s.V * 4;
    }
    TMVB_TRY(filt_push_kq(&s, h->d_kappa, h->d_kq, h->eta));
    if (tau) {
        TMVB_CHECK_ARG(s.corpus_set, "set_corpus must precede the upload of tau");
        const size_t need = (size_t)std::max<int64_t>(s.nnz, 1);
        if (need > h->tau_cap) {
            cudaFree(h->d_tau);
            cudaFree(h->d_tau_old);
            h->d_tau = h->d_tau_old = nullptr;
            h->tau_cap = 0;
            TMVB_CUDA(cudaMalloc((void **)&h->d_tau, need * 4));
            TMVB_CUDA(cudaMalloc((void **)&h->d_tau_old, need * 4));
            h->tau_cap = need;
        }
        int terr = 0;
        TMVB_TRY(filt_tau_upload(&s, tau, h->d_tau, &terr));
        if (s.nnz > 0) TMVB_CUDA(cudaMemcpyAsync(h->d_tau_old, h->d_tau, (size_t)s.nnz * 4, cudaMemcpyDeviceToDevice, s.stream));
        if (terr) return fail(-5, "tau must contain probabilities.");
        h->filt_set = true;
    }
    return 0;
}

int tmvb_fctm_download(tmvb_ctm_t h, float *kappa, float *kappa_old, float *tau, float *tau_old)
{
    TMVB_CHECK_ARG(h != nullptr && h->filtered, "not a filtered-CTM handle");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    if (kappa && s.V > 0) TMVB_CUDA(cudaMemcpyAsync(kappa, h->d_kappa, s.V * 4, cudaMemcpyDeviceToHost, s.stream));
    if (kappa_old && s.V > 0) TMVB_CUDA(cudaMemcpyAsync(kappa_old, h->d_kappa_old, s.V * 4, cudaMemcpyDeviceToHost, s.stream));
    s.st.d2h_bytes += ((kappa ? 1 : 0) + (kappa_old ? 1 : 0)) * s.V * 4;
    if (tau && h->d_tau) TMVB_TRY(filt_tau_download(&s, h->d_tau, tau));
    if (tau_old && h->d_tau_old) TMVB_TRY(filt_tau_download(&s, h->d_tau_old, tau_old));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    return 0;
}

/* the update_kappa! statistics: summed over ranks next to the buffers of tmvb_ctm_reduce_buffers */
int tmvb_fctm_reduce_buffers(tmvb_ctm_t h, void **kstats, int64_t *n_kstats)
{
    TMVB_CHECK_ARG(h != nullptr && h->filtered, "not a filtered-CTM handle");
    if (kstats) *kstats = h->d_kstats;
    if (n_kstats) *n_kstats = h->s.V;
    return 0;
}

/* ---- multi-GPU: the statistics summed over the ranks by ONE kernel over CUDA-IPC peer memory (tmvb_peer.cu) instead of one NCCL
 * all-reduce per buffer.  Handshake as for gpuLDA: export -> all-gather the blobs over any transport -> connect; then
 * tmvb_ctm_peer_reduce(h) between estep and mstep on every rank. ---- */
int tmvb_ctm_comm_export(tmvb_ctm_t h, void *blob, int64_t blob_bytes)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CHECK_ARG(blob_bytes >= TMVB_COMM_BLOB_BYTES, "blob must hold TMVB_COMM_BLOB_BYTES");
    TMVB_CUDA(cudaSetDevice(h->s.device));
    return peer_export(&h->comm, ctm_peer_bufs(h), blob, (size_t)blob_bytes);
}

int tmvb_ctm_comm_connect(tmvb_ctm_t h, int rank, int world, const void *blobs, int64_t blob_bytes)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CUDA(cudaSetDevice(h->s.device));
    TMVB_CUDA(cudaStreamSynchronize(h->s.stream));
    return comm_connect(&h->comm, rank, world, blobs, (size_t)blob_bytes);
}

int tmvb_ctm_peer_reduce(tmvb_ctm_t h)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CUDA(cudaSetDevice(h->s.device));
    TMVB_TRY(peer_allreduce(&h->comm, ctm_peer_bufs(h), h->s.stream, h->s.n_sm));
    h->s.st.kernel_launches++;
    return 0;
}

int tmvb_ctm_get_stats(tmvb_ctm_t h, tmvb_stats *out)
{
    TMVB_CHECK_ARG(h && out, "NULL argument");
    return shard_get_stats(&h->s, h->d_small + 1, out);
}

}  // extern "C"
