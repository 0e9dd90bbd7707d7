// tmvb_estep.cuh -- device building blocks shared by the three E-step kernels (LDA, CTM, CTPF).
//
// All three models update, per document, a K-vector from N_d "tokens" whose responsibilities are a
// softmax over topics of (table row of the token) x (per-document vector e):
//     phi_ni = (eps + T[i, w_n] e_i) / s_n ,   s_n = K eps + sum_i T[i, w_n] e_i
// LDA: T = beta, e = exp(Elogtheta) (LDA.jl:150-154);  CTM: T = beta, e = exp(lambda - max), eps = 0
// (CTM.jl:175-178);  CTPF: T = exp(psi(alef)) resp. exp(psi(he)), e from the Gamma shapes/rates
// (CTPF.jl:327-337).  The quantity every sweep needs is (phi * counts)_i = e_i g_i + eps sum_n t_n with
// t_n = c_n / s_n and g_i = sum_n T[i, w_n] t_n: two FMA passes over a shared-memory tile of table rows.
//
// Thread mapping: ONE WARP PER DOCUMENT.
//   token phase -- lane (ts = lane / LPT, kl = lane % LPT) owns token stream ts (S = 32 / LPT streams) and the
//                  16-byte chunks q = kl + LPT*m (m < CPL) of every row it visits (topics 4q..4q+3, LDS.128).
//   K phase     -- lane l owns topics i = l + 32 r (r < R).
// The layouts meet in shared memory: per-stream partial K-vectors are written as float4 chunks (gs) and
// summed by the owner lanes; the per-document vector e travels back through e_s.  The row stride RS of
// tile/gs is padded so that RS/4 = LPT (mod 2 LPT) for LPT < 8, which makes every LDS.128/STS.128 phase
// bank-conflict free.
#pragma once

#include "tmvb_common.cuh"

namespace tmvb {

// rounds of the sweep pass kept in flight per warp (independent LDS.128 -> FFMA -> shuffle chains)
constexpr int kSweepUnroll = 2;

#ifdef __CUDACC__

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

// What one pass over a document's tokens needs.
struct TokArgs {
    const float *tile;     // shared: [cap][RS] staged table rows
    const float *cnt_s;    // shared: [cap] counts
    const int *term_s;     // shared: [cap] row ids
    const float *gtable;   // global: [rows][K_ld] table (overflow rows are read from here)
    const int *gterms;     // global: this document's row ids   (already offset to the document)
    const float *gcounts;  // global: this document's counts
    float *stats;          // global: [rows][K_ld] scatter target (final pass)
    int Nd, cap, rounds, K, K_ld, RS, dbg;
    int r0, rstep;         // this warp takes rounds r0, r0 + rstep, ... (rstep = warps cooperating on the document)
};

// Sweep pass: s_n, t_n = c_n / s_n, g += T t.  EPS selects the reference's "@positive" epsilon (LDA) or none.
template <int LPT, int CPL, bool OVF, bool EPS, int UNR = kSweepUnroll>
__device__ __forceinline__ void tok_sweep(const TokArgs &a, int ts, int kl, const float4 (&e)[CPL], float4 (&g)[CPL], float &tsum)
{
    constexpr int S = 32 / LPT;
    const int CH = a.K_ld >> 2;
    const float Keps = EPS ? (float)a.K * TMVB_EPS : 0.0f;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll UNR
    for (int r = a.r0; r < a.rounds; r += a.rstep) {
        const int n = r * S + ts;
        const bool ok = n < a.Nd;
        float4 b[CPL];
        float c = 0.0f;
        if (!OVF || n < a.cap) {
            const int nn = ok ? n : 0;
            const float4 *row = reinterpret_cast<const float4 *>(a.tile + nn * a.RS) + kl;
#pragma unroll
            for (int m = 0; m < CPL; m++) b[m] = (m < CPL - 1 || kl + LPT * m < CH) ? row[LPT * m] : zero4;
            if (ok) c = a.cnt_s[nn];
        } else {
            const int nn = ok ? n : 0;
            const float4 *row = reinterpret_cast<const float4 *>(a.gtable + (size_t)a.gterms[nn] * a.K_ld) + kl;
#pragma unroll
            for (int m = 0; m < CPL; m++) b[m] = (m < CPL - 1 || kl + LPT * m < CH) ? __ldg(row + LPT * m) : zero4;
            if (ok) c = a.gcounts[nn];
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
        const float t = ok ? __fdividef(c, s) : 0.0f;
#pragma unroll
        for (int m = 0; m < CPL; m++) {
            g[m].x = fmaf(b[m].x, t, g[m].x);
            g[m].y = fmaf(b[m].y, t, g[m].y);
            g[m].z = fmaf(b[m].z, t, g[m].z);
            g[m].w = fmaf(b[m].w, t, g[m].w);
        }
        tsum += t;
    }
}

// Final pass: scatter c_n phi_ni = t_n (eps + T e_i) into stats with 16-byte vector reductions
// (REDG.E.ADD.F32x4) and, when ELBO, accumulate sum_n c_n H(phi_n) = sum_n [c_n ln s_n - sum_i c_n phi_ni ln u_ni].
template <int LPT, int CPL, bool OVF, bool EPS, bool ELBO>
__device__ __forceinline__ void tok_final(const TokArgs &a, int ts, int kl, const float4 (&e)[CPL], float &ent)
{
    constexpr int S = 32 / LPT;
    const int CH = a.K_ld >> 2;
    const float Keps = EPS ? (float)a.K * TMVB_EPS : 0.0f;
    const float eps = EPS ? TMVB_EPS : 0.0f;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = a.r0; r < a.rounds; r += a.rstep) {
        const int n = r * S + ts;
        const bool ok = n < a.Nd;
        float4 b[CPL];
        float c = 0.0f;
        int term = 0;
        if (!OVF || n < a.cap) {
            const int nn = ok ? n : 0;
            const float4 *row = reinterpret_cast<const float4 *>(a.tile + nn * a.RS) + kl;
#pragma unroll
            for (int m = 0; m < CPL; m++) b[m] = (m < CPL - 1 || kl + LPT * m < CH) ? row[LPT * m] : zero4;
            if (ok) c = a.cnt_s[nn];
            term = a.term_s[nn];
        } else {
            const int nn = ok ? n : 0;
            term = a.gterms[nn];
            const float4 *row = reinterpret_cast<const float4 *>(a.gtable + (size_t)term * a.K_ld) + kl;
#pragma unroll
            for (int m = 0; m < CPL; m++) b[m] = (m < CPL - 1 || kl + LPT * m < CH) ? __ldg(row + LPT * m) : zero4;
            if (ok) c = a.gcounts[nn];
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
        if (ok) {
            const float t = __fdividef(c, s);
            float *srow = a.stats + (size_t)term * a.K_ld + 4 * kl;
            float acc = 0.0f;
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                const int i0 = 4 * (kl + LPT * m);
                if (i0 < a.K) {
                    // pad topics (i >= K) carry T = e = 0: they receive t*eps, which the M-step ignores
                    const float ux = fmaf(b[m].x, e[m].x, eps), uy = fmaf(b[m].y, e[m].y, eps);
                    const float uz = fmaf(b[m].z, e[m].z, eps), uw = fmaf(b[m].w, e[m].w, eps);
                    if (!(a.dbg & 1)) red_add_v4(srow + 4 * LPT * m, t * ux, t * uy, t * uz, t * uw);
                    if (ELBO) {
                        if (EPS || ux > 0.f) acc = fmaf(t * ux, __logf(ux), acc);
                        if (i0 + 1 < a.K && (EPS || uy > 0.f)) acc = fmaf(t * uy, __logf(uy), acc);
                        if (i0 + 2 < a.K && (EPS || uz > 0.f)) acc = fmaf(t * uz, __logf(uz), acc);
                        if (i0 + 3 < a.K && (EPS || uw > 0.f)) acc = fmaf(t * uw, __logf(uw), acc);
                    }
                }
            }
            if (ELBO) ent += ((kl == 0) ? c * __logf(s) : 0.0f) - acc;
        }
    }
}

// Final pass for a 2K-way softmax that shares one table row: xi_r = [T ea ; T eb] / s_r, s_r = sum_i T_i (ea_i + eb_i)
// (CTPF.jl:334-337).  Scatters rating_r (xi_a + xi_b) (CTPF.jl:274-277) and accumulates rating_r H(xi_r).
template <int LPT, int CPL, bool OVF, bool ELBO>
__device__ __forceinline__ void tok_final2(const TokArgs &a, int ts, int kl, const float4 (&ea)[CPL], const float4 (&eb)[CPL], float &ent)
{
    constexpr int S = 32 / LPT;
    const int CH = a.K_ld >> 2;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = a.r0; r < a.rounds; r += a.rstep) {
        const int n = r * S + ts;
        const bool ok = n < a.Nd;
        float4 b[CPL];
        float c = 0.0f;
        int term = 0;
        const int nn = ok ? n : 0;
        if (!OVF || n < a.cap) {
            const float4 *row = reinterpret_cast<const float4 *>(a.tile + nn * a.RS) + kl;
#pragma unroll
            for (int m = 0; m < CPL; m++) b[m] = (m < CPL - 1 || kl + LPT * m < CH) ? row[LPT * m] : zero4;
            if (ok) c = a.cnt_s[nn];
            term = a.term_s[nn];
        } else {
            term = a.gterms[nn];
            const float4 *row = reinterpret_cast<const float4 *>(a.gtable + (size_t)term * a.K_ld) + kl;
#pragma unroll
            for (int m = 0; m < CPL; m++) b[m] = (m < CPL - 1 || kl + LPT * m < CH) ? __ldg(row + LPT * m) : zero4;
            if (ok) c = a.gcounts[nn];
        }
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int m = 0; m < CPL; m++) {
            s0 = fmaf(b[m].x, ea[m].x + eb[m].x, s0);
            s1 = fmaf(b[m].y, ea[m].y + eb[m].y, s1);
            s0 = fmaf(b[m].z, ea[m].z + eb[m].z, s0);
            s1 = fmaf(b[m].w, ea[m].w + eb[m].w, s1);
        }
        const float s = group_sum<LPT>(s0 + s1);
        if (ok) {
            const float t = __fdividef(c, s);
            float *srow = a.stats + (size_t)term * a.K_ld + 4 * kl;
            float acc = 0.0f;
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                const int i0 = 4 * (kl + LPT * m);
                if (i0 < a.K) {
                    const float ax = b[m].x * ea[m].x, ay = b[m].y * ea[m].y, az = b[m].z * ea[m].z, aw = b[m].w * ea[m].w;
                    const float bx = b[m].x * eb[m].x, by = b[m].y * eb[m].y, bz = b[m].z * eb[m].z, bw = b[m].w * eb[m].w;
                    if (!(a.dbg & 1)) red_add_v4(srow + 4 * LPT * m, t * (ax + bx), t * (ay + by), t * (az + bz), t * (aw + bw));
                    if (ELBO) {
                        if (ax > 0.f) acc = fmaf(t * ax, __logf(ax), acc);
                        if (bx > 0.f) acc = fmaf(t * bx, __logf(bx), acc);
                        if (i0 + 1 < a.K && ay > 0.f) acc = fmaf(t * ay, __logf(ay), acc);
                        if (i0 + 1 < a.K && by > 0.f) acc = fmaf(t * by, __logf(by), acc);
                        if (i0 + 2 < a.K && az > 0.f) acc = fmaf(t * az, __logf(az), acc);
                        if (i0 + 2 < a.K && bz > 0.f) acc = fmaf(t * bz, __logf(bz), acc);
                        if (i0 + 3 < a.K && aw > 0.f) acc = fmaf(t * aw, __logf(aw), acc);
                        if (i0 + 3 < a.K && bw > 0.f) acc = fmaf(t * bw, __logf(bw), acc);
                    }
                }
            }
            if (ELBO) ent += ((kl == 0) ? c * __logf(s) : 0.0f) - acc;
        }
    }
}

// owner-lane sum of the S per-stream partials of topic i (4 independent chains)
template <int S>
__device__ __forceinline__ float owner_sum(const float *gs, int RS, int i)
{
    float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f, g3 = 0.0f;
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
    return (g0 + g1) + (g2 + g3);
}

// Stage `ns` table rows (ids in term_s) into the tile: one TMA bulk copy per row, completion on `mbar`
// (stage_bulk) or 16-byte cp.async per lane.  Caller waits with stage_wait().
__device__ __forceinline__ void stage_rows(float *tile, const int *term_s, const float *gtable, int ns, int K_ld, int RS, int lane,
                                           unsigned long long *mbar, int stage_bulk)
{
    if (stage_bulk) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(mbar, (unsigned)(ns * K_ld * 4));
        __syncwarp();
        for (int n = lane; n < ns; n += 32) bulk_g2s(tile + n * RS, gtable + (size_t)term_s[n] * K_ld, (unsigned)(K_ld * 4), mbar);
    } else {
        const int CH = K_ld >> 2;
        __syncwarp();
        for (int c = lane; c < ns * CH; c += 32) {
            const int n = c / CH, q = c - n * CH;
            cp_async16(tile + n * RS + 4 * q, gtable + (size_t)term_s[n] * K_ld + 4 * q);
        }
        cp_async_commit();
    }
}
__device__ __forceinline__ void stage_wait(unsigned long long *mbar, unsigned &phase, int stage_bulk)
{
    if (stage_bulk) {
        mbar_wait(mbar, phase);
        phase ^= 1u;
    } else {
        cp_async_wait_all();
    }
    __syncwarp();
}

#endif  // __CUDACC__

// ---------------------------------------------------------------- host: lane layouts ----------
struct LaneLayout {
    int lpt, cpl;
};
// the (LPT, CPL) pairs pick_layout can select for K <= 256
static const LaneLayout kLaneLayouts[] = {{1, 4}, {1, 8}, {2, 1}, {2, 3}, {2, 5}, {2, 6}, {2, 7}, {2, 8},
                                          {4, 5}, {4, 6}, {4, 7}, {4, 8}, {8, 5}, {8, 6}, {8, 7}, {8, 8}};
constexpr int kNumLaneLayouts = sizeof(kLaneLayouts) / sizeof(kLaneLayouts[0]);

// expands X(LPT, CPL) for every entry of kLaneLayouts, in the same order
#define TMVB_FOR_EACH_LAYOUT(X) \
    X(1, 4) X(1, 8) X(2, 1) X(2, 3) X(2, 5) X(2, 6) X(2, 7) X(2, 8) X(4, 5) X(4, 6) X(4, 7) X(4, 8) X(8, 5) X(8, 6) X(8, 7) X(8, 8)

// row stride (floats) of the shared-memory tile for CH 16-byte chunks per row
int row_stride(int CH, int lpt);
// index into kLaneLayouts minimising (estimated warp-instructions per token) x sqrt(shared-memory inflation); -1 if none
int pick_layout(int K_ld, int *RS_out);

}  // namespace tmvb
