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

// A staged table row as this lane sees it: CPL 16-byte chunks, each two packed fp32 pairs (.x = topics 4q, 4q+1; .y = 4q+2, 4q+3)
// so that the FMA passes issue as FFMA2 (fma.rn.f32x2): half the floating-point issue slots of scalar FFMA.
template <int LPT, int CPL, bool OVF>
__device__ __forceinline__ void tok_load(const TokArgs &a, int n, bool ok, int kl, int CH, ulonglong2 (&b)[CPL], float &c, int &term, bool want_term)
{
    const ulonglong2 zero = make_ulonglong2(0ull, 0ull);
    const int nn = ok ? n : 0;
    c = 0.0f;
    term = 0;
    if (!OVF || n < a.cap) {
        const ulonglong2 *row = reinterpret_cast<const ulonglong2 *>(a.tile + nn * a.RS) + kl;
#pragma unroll
        for (int m = 0; m < CPL; m++) b[m] = (m < CPL - 1 || kl + LPT * m < CH) ? row[LPT * m] : zero;
        if (ok) c = a.cnt_s[nn];
        if (want_term) term = a.term_s[nn];
    } else {
        term = a.gterms[nn];
        const ulonglong2 *row = reinterpret_cast<const ulonglong2 *>(a.gtable + (size_t)term * a.K_ld) + kl;
#pragma unroll
        for (int m = 0; m < CPL; m++) b[m] = (m < CPL - 1 || kl + LPT * m < CH) ? __ldg(row + LPT * m) : zero;
        if (ok) c = a.gcounts[nn];
    }
}

// s = sum_i T_i e_i over this lane's chunks (two or four independent FFMA2 chains), then across the LPT lanes of the token
template <int LPT, int CPL>
__device__ __forceinline__ float tok_dot(const ulonglong2 (&b)[CPL], const f32x2 (&e01)[CPL], const f32x2 (&e23)[CPL])
{
    f32x2 sa = 0ull, sb = 0ull, sc = 0ull, sd = 0ull;
#pragma unroll
    for (int m = 0; m < CPL; m++) {
        if (CPL >= 4 && (m & 1)) {
            sc = fma2(b[m].x, e01[m], sc);
            sd = fma2(b[m].y, e23[m], sd);
        } else {
            sa = fma2(b[m].x, e01[m], sa);
            sb = fma2(b[m].y, e23[m], sb);
        }
    }
    if (CPL >= 4) {
        sa = add2(sa, sc);
        sb = add2(sb, sd);
    }
    return group_sum<LPT>(hsum2(add2(sa, sb)));
}

// Sweep pass: s_n, t_n = c_n / s_n, g += T t.  EPS selects the reference's "@positive" epsilon (LDA) or none.
template <int LPT, int CPL, bool OVF, bool EPS, int UNR = kSweepUnroll>
__device__ __forceinline__ void tok_sweep(const TokArgs &a, int ts, int kl, const float4 (&e)[CPL], float4 (&g)[CPL], float &tsum)
{
    constexpr int S = 32 / LPT;
    const int CH = a.K_ld >> 2;
    const float Keps = EPS ? (float)a.K * TMVB_EPS : 0.0f;
    f32x2 e01[CPL], e23[CPL], g01[CPL], g23[CPL];
#pragma unroll
    for (int m = 0; m < CPL; m++) {
        e01[m] = pk2(e[m].x, e[m].y);
        e23[m] = pk2(e[m].z, e[m].w);
        g01[m] = pk2(g[m].x, g[m].y);
        g23[m] = pk2(g[m].z, g[m].w);
    }
#pragma unroll UNR
    for (int r = a.r0; r < a.rounds; r += a.rstep) {
        const int n = r * S + ts;
        const bool ok = n < a.Nd;
        ulonglong2 b[CPL];
        float c;
        int term;
        tok_load<LPT, CPL, OVF>(a, n, ok, kl, CH, b, c, term, false);
        const float s = tok_dot<LPT, CPL>(b, e01, e23) + Keps;
        const float t = ok ? __fdividef(c, s) : 0.0f;
        const f32x2 t2 = pk2(t, t);
#pragma unroll
        for (int m = 0; m < CPL; m++) {
            g01[m] = fma2(b[m].x, t2, g01[m]);
            g23[m] = fma2(b[m].y, t2, g23[m]);
        }
        tsum += t;
    }
#pragma unroll
    for (int m = 0; m < CPL; m++) {
        unpk2(g01[m], g[m].x, g[m].y);
        unpk2(g23[m], g[m].z, g[m].w);
    }
}

// Final pass: scatter c_n phi_ni = t_n (eps + T e_i) into stats with 16-byte vector reductions
// (REDG.E.ADD.F32x4).  ELBO = 1: also accumulate sum_n c_n H(phi_n) = sum_n [c_n ln s_n - sum_i c_n phi_ni ln u_ni];
// ELBO = 2: only sum_n c_n ln s_n -- the caller supplies sum_{n,i} c_n phi_ni ln u_ni in closed form from K- and
// K x V-sized quantities (ln u_ni = ln T_i,w + ln e_i), so no per-(token, topic) logarithm is evaluated.
template <int LPT, int CPL, bool OVF, bool EPS, int ELBO>
__device__ __forceinline__ void tok_final(const TokArgs &a, int ts, int kl, const float4 (&e)[CPL], float &ent)
{
    constexpr int S = 32 / LPT;
    const int CH = a.K_ld >> 2;
    const float Keps = EPS ? (float)a.K * TMVB_EPS : 0.0f;
    const float eps = EPS ? TMVB_EPS : 0.0f;
    const f32x2 eps2 = pk2(eps, eps);
    f32x2 e01[CPL], e23[CPL];
#pragma unroll
    for (int m = 0; m < CPL; m++) {
        e01[m] = pk2(e[m].x, e[m].y);
        e23[m] = pk2(e[m].z, e[m].w);
    }
    for (int r = a.r0; r < a.rounds; r += a.rstep) {
        const int n = r * S + ts;
        const bool ok = n < a.Nd;
        ulonglong2 b[CPL];
        float c;
        int term;
        tok_load<LPT, CPL, OVF>(a, n, ok, kl, CH, b, c, term, true);
        const float s = tok_dot<LPT, CPL>(b, e01, e23) + Keps;
        if (ok) {
            const float t = __fdividef(c, s);
            const f32x2 t2 = pk2(t, t);
            float *srow = a.stats + (size_t)term * a.K_ld + 4 * kl;
            float acc = 0.0f;
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                const int i0 = 4 * (kl + LPT * m);
                if (i0 < a.K) {
                    // pad topics (i >= K) carry T = e = 0: they receive t*eps, which the M-step ignores
                    const f32x2 u01 = fma2(b[m].x, e01[m], eps2), u23 = fma2(b[m].y, e23[m], eps2);
                    float px, py, pz, pw;
                    unpk2(mul2(t2, u01), px, py);
                    unpk2(mul2(t2, u23), pz, pw);
                    if (!(a.dbg & 1)) red_add_v4(srow + 4 * LPT * m, px, py, pz, pw);
                    if (ELBO == 1) {
                        float ux, uy, uz, uw;
                        unpk2(u01, ux, uy);
                        unpk2(u23, uz, uw);
                        if (EPS || ux > 0.f) acc = fmaf(px, __logf(ux), acc);
                        if (i0 + 1 < a.K && (EPS || uy > 0.f)) acc = fmaf(py, __logf(uy), acc);
                        if (i0 + 2 < a.K && (EPS || uz > 0.f)) acc = fmaf(pz, __logf(uz), acc);
                        if (i0 + 3 < a.K && (EPS || uw > 0.f)) acc = fmaf(pw, __logf(uw), acc);
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
    f32x2 ea01[CPL], ea23[CPL], eb01[CPL], eb23[CPL], es01[CPL], es23[CPL];
#pragma unroll
    for (int m = 0; m < CPL; m++) {
        ea01[m] = pk2(ea[m].x, ea[m].y);
        ea23[m] = pk2(ea[m].z, ea[m].w);
        eb01[m] = pk2(eb[m].x, eb[m].y);
        eb23[m] = pk2(eb[m].z, eb[m].w);
        es01[m] = add2(ea01[m], eb01[m]);
        es23[m] = add2(ea23[m], eb23[m]);
    }
    for (int r = a.r0; r < a.rounds; r += a.rstep) {
        const int n = r * S + ts;
        const bool ok = n < a.Nd;
        ulonglong2 b[CPL];
        float c;
        int term;
        tok_load<LPT, CPL, OVF>(a, n, ok, kl, CH, b, c, term, true);
        const float s = tok_dot<LPT, CPL>(b, es01, es23);
        if (ok) {
            const float t = __fdividef(c, s);
            const f32x2 t2 = pk2(t, t);
            float *srow = a.stats + (size_t)term * a.K_ld + 4 * kl;
            float acc = 0.0f;
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                const int i0 = 4 * (kl + LPT * m);
                if (i0 < a.K) {
                    float px, py, pz, pw;
                    unpk2(mul2(t2, mul2(b[m].x, es01[m])), px, py);
                    unpk2(mul2(t2, mul2(b[m].y, es23[m])), pz, pw);
                    if (!(a.dbg & 1)) red_add_v4(srow + 4 * LPT * m, px, py, pz, pw);
                    if (ELBO) {
                        float ax, ay, az, aw, bx, by, bz, bw;
                        unpk2(mul2(b[m].x, ea01[m]), ax, ay);
                        unpk2(mul2(b[m].y, ea23[m]), az, aw);
                        unpk2(mul2(b[m].x, eb01[m]), bx, by);
                        unpk2(mul2(b[m].y, eb23[m]), bz, bw);
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

// ---------------------------------------------------------------- register-resident documents ----------
// For documents of at most W * NR * S tokens the K x N_d slab of the table never touches shared memory: warp w of the
// W warps that share a document keeps rounds j = 0..NR-1 (token n = ((j W + w) S + ts)) in registers for all sweeps
// (NR * CPL 16-byte chunks per lane), loaded once with LDG.128 straight from L2.  The loops over j are fully unrolled,
// so the NR rounds of a sweep are independent instruction streams the scheduler can interleave.
template <int LPT, int CPL, int NR>
struct RegDoc {
    ulonglong2 b[NR][CPL];
    float c[NR];
    int term[NR];
};

template <int LPT, int CPL, int W, int NR>
__device__ __forceinline__ void reg_load(RegDoc<LPT, CPL, NR> &rd, const float *__restrict__ gtable, const int *__restrict__ gterms,
                                         const float *__restrict__ gcounts, int Nd, int K_ld, int warp, int ts, int kl, int dbg)
{
    constexpr int S = 32 / LPT;
    const int CH = K_ld >> 2;
    const ulonglong2 zero = make_ulonglong2(0ull, 0ull);
#pragma unroll
    for (int j = 0; j < NR; j++) {
        const int n = (j * W + warp) * S + ts;
        const bool ok = n < Nd;
        rd.term[j] = ok ? __ldg(gterms + n) : 0;
        rd.c[j] = ok ? __ldg(gcounts + n) : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < NR; j++) {
        const int n = (j * W + warp) * S + ts;
        // dbg bit 2 (developer probe): every token reads row 0 -- isolates the cost of the random row gather
        const ulonglong2 *row = reinterpret_cast<const ulonglong2 *>(gtable + ((dbg & 4) ? (size_t)0 : (size_t)rd.term[j] * K_ld)) + kl;
#pragma unroll
        for (int m = 0; m < CPL; m++) rd.b[j][m] = (n < Nd && (m < CPL - 1 || kl + LPT * m < CH)) ? __ldg(row + LPT * m) : zero;
    }
}

// one sweep over the register-resident rounds: s_n, t_n = c_n / s_n, g += T t (cf. tok_sweep)
template <int LPT, int CPL, int NR, bool EPS>
__device__ __forceinline__ void reg_sweep(const RegDoc<LPT, CPL, NR> &rd, int K, const f32x2 (&e01)[CPL], const f32x2 (&e23)[CPL],
                                          f32x2 (&g01)[CPL], f32x2 (&g23)[CPL], float &tsum)
{
    const float Keps = EPS ? (float)K * TMVB_EPS : 0.0f;
    float t[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) {
        const float s = tok_dot<LPT, CPL>(rd.b[j], e01, e23) + Keps;
        t[j] = (rd.c[j] > 0.0f) ? (EPS ? fast_div_pos(rd.c[j], s) : __fdividef(rd.c[j], s)) : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < NR; j++) {
        const f32x2 t2 = pk2(t[j], t[j]);
#pragma unroll
        for (int m = 0; m < CPL; m++) {
            g01[m] = fma2(rd.b[j][m].x, t2, g01[m]);
            g23[m] = fma2(rd.b[j][m].y, t2, g23[m]);
        }
        tsum += t[j];
    }
}

// final pass over the register-resident rounds: scatter t_n (eps + T e) into stats (REDG.E.ADD.F32x4) and, when ELBO,
// accumulate sum_n c_n ln s_n (the "ELBO = 2" form of tok_final)
template <int LPT, int CPL, int NR, bool EPS, bool ELBO>
__device__ __forceinline__ void reg_final(const RegDoc<LPT, CPL, NR> &rd, float *__restrict__ stats, int K, int K_ld, int kl,
                                          const f32x2 (&e01)[CPL], const f32x2 (&e23)[CPL], float &ent, int dbg)
{
    const float Keps = EPS ? (float)K * TMVB_EPS : 0.0f;
    const float eps = EPS ? TMVB_EPS : 0.0f;
    const f32x2 eps2 = pk2(eps, eps);
#pragma unroll
    for (int j = 0; j < NR; j++) {
        const float s = tok_dot<LPT, CPL>(rd.b[j], e01, e23) + Keps;
        if (rd.c[j] > 0.0f) {
            const float t = EPS ? fast_div_pos(rd.c[j], s) : __fdividef(rd.c[j], s);
            const f32x2 t2 = pk2(t, t);
            float *srow = stats + (size_t)rd.term[j] * K_ld + 4 * kl;
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                if (4 * (kl + LPT * m) < K) {
                    float px, py, pz, pw;
                    unpk2(mul2(t2, fma2(rd.b[j][m].x, e01[m], eps2)), px, py);
                    unpk2(mul2(t2, fma2(rd.b[j][m].y, e23[m], eps2)), pz, pw);
                    if (!(dbg & 1)) red_add_v4(srow + 4 * LPT * m, px, py, pz, pw);
                }
            }
            if (ELBO && kl == 0) ent += rd.c[j] * __logf(s);
        }
    }
}

// ---------------------------------------------------------------- bare MUFU forms for the per-sweep hot path ----------
// __fdividef / __logf / __expf carry range guards (denormal scaling: 4-6 extra instructions each); the arguments of the
// K phase are normal numbers by construction (gamma >= alpha + eps, products of (x + k), exp of a non-positive number whose
// flush to zero below 2^-126 is harmless), so the bare approximations are used: one MUFU and at most one multiply each.
__device__ __forceinline__ float rcp_ftz(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_ftz(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_ftz(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// psi(x) for two arguments in packed fp32, branch-free: the shift-by-6 rational recurrence of psi_lgamma (one reciprocal
// instead of six) and the 3-term asymptotic series, selected per argument.  Same arithmetic as psi_pair up to the guards.
__device__ __forceinline__ void psi_pair_fast(float x0, float x1, float &p0, float &p1)
{
    const f32x2 x = pk2(x0, x1);
    f32x2 P = x, D = pk2(1.0f, 1.0f);
#pragma unroll
    for (int k = 1; k < 6; k++) {
        const f32x2 f = add2(x, pk2((float)k, (float)k));
        D = fma2(D, f, P);
        P = mul2(P, f);
    }
    float P0, P1, D0, D1;
    unpk2(P, P0, P1);
    unpk2(D, D0, D1);
    const bool lo0 = x0 < 6.0f, lo1 = x1 < 6.0f;
    const float y0 = lo0 ? x0 + 6.0f : x0, y1 = lo1 ? x1 + 6.0f : x1;
    const float c0 = lo0 ? D0 * rcp_ftz(P0) : 0.0f, c1 = lo1 ? D1 * rcp_ftz(P1) : 0.0f;
    const f32x2 t = pk2(rcp_ftz(y0), rcp_ftz(y1)), nt2 = mul2(mul2(t, t), pk2(-1.0f, -1.0f));
    // psi = ln y - t/2 - t2 (1/12 - t2 (1/120 - t2/252)) - corr
    const f32x2 in1 = fma2(nt2, pk2(3.9682539683e-3f, 3.9682539683e-3f), pk2(8.3333333333e-3f, 8.3333333333e-3f));
    const f32x2 in2 = fma2(nt2, in1, pk2(8.3333333333e-2f, 8.3333333333e-2f));
    f32x2 r = fma2(t, pk2(-0.5f, -0.5f), pk2(fmaf(lg2_ftz(y0), 0.69314718056f, -c0), fmaf(lg2_ftz(y1), 0.69314718056f, -c1)));
    r = fma2(nt2, in2, r);
    unpk2(r, p0, p1);
}

// s = init + sum_i T_i e_i over this lane's chunks in two FFMA2 chains (FFMA2 issues every other cycle, so two chains of a
// round plus the other rounds in flight cover its latency), then across the LPT lanes of the token.  `init` carries K eps in
// one lane of the token's group (the "@positive" epsilon of LDA.jl:150-154 summed over topics).
#ifndef TMVB_V_DOT4
#define TMVB_V_DOT4 0   // 1: four FFMA2 chains per dot product instead of two (A/B switch, tools/build_variants.sh)
#endif
template <int LPT, int CPL>
__device__ __forceinline__ float tok_dot2(const ulonglong2 (&b)[CPL], const f32x2 (&e01)[CPL], const f32x2 (&e23)[CPL], f32x2 init)
{
#if TMVB_V_DOT4
    f32x2 sa = init, sb = 0ull, sc = 0ull, sd = 0ull;
#pragma unroll
    for (int m = 0; m < CPL; m++) {
        if (m & 1) {
            sc = fma2(b[m].x, e01[m], sc);
            sd = fma2(b[m].y, e23[m], sd);
        } else {
            sa = fma2(b[m].x, e01[m], sa);
            sb = fma2(b[m].y, e23[m], sb);
        }
    }
    return group_sum<LPT>(hsum2(add2(add2(sa, sc), add2(sb, sd))));
#else
    f32x2 sa = init, sb = 0ull;
#pragma unroll
    for (int m = 0; m < CPL; m++) {
        sa = fma2(b[m].x, e01[m], sa);
        sb = fma2(b[m].y, e23[m], sb);
    }
    return group_sum<LPT>(hsum2(add2(sa, sb)));
#endif
}

// ---------------------------------------------------------------- hybrid documents: registers + shared-memory tile ----------
// The register-resident rounds of a document (RegDoc) are complemented by a shared-memory tile for the tokens beyond
// W * NR * S: tile round q (tokens q S + ts of the tile) belongs to warp q % W.  Registers hold what the register file has
// room for at the target occupancy; the tile holds the rest, so one kernel serves every document length without
// falling off a cliff (a pure tile kernel is bound by the shared-memory pipe, a pure register kernel by occupancy).
//
// Sweep over the register rounds that also returns s_n (kept for the scatter pass: the last sweep's s_n is the
// normaliser of the phi that update_beta! scatters, LDA.jl:129-132, so the final pass needs no second dot product).
template <int LPT, int CPL, int NR, bool EPS>
__device__ __forceinline__ void reg_sweep_keep(const RegDoc<LPT, CPL, NR> &rd, f32x2 keps_init, const f32x2 (&e01)[CPL], const f32x2 (&e23)[CPL],
                                               f32x2 (&g01)[CPL], f32x2 (&g23)[CPL], float &tsum, float (&s_keep)[NR])
{
    static_assert(EPS, "without the epsilon a padded token (all-zero row) would divide by zero");
    float t[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) {
        const float s = tok_dot2<LPT, CPL>(rd.b[j], e01, e23, keps_init);
        s_keep[j] = s;
        t[j] = rd.c[j] * rcp_ftz(s);   // padded tokens: c = 0, s = K eps > 0
    }
#pragma unroll
    for (int j = 0; j < NR; j++) {
        const f32x2 t2 = pk2(t[j], t[j]);
#pragma unroll
        for (int m = 0; m < CPL; m++) {
            g01[m] = fma2(rd.b[j][m].x, t2, g01[m]);
            g23[m] = fma2(rd.b[j][m].y, t2, g23[m]);
        }
        tsum += t[j];
    }
}

// scatter pass over the register rounds from the kept normalisers (cf. reg_final)
template <int LPT, int CPL, int NR, bool EPS, bool ELBO>
__device__ __forceinline__ void reg_final_keep(const RegDoc<LPT, CPL, NR> &rd, float *__restrict__ stats, int K, int K_ld, int kl,
                                               const f32x2 (&e01)[CPL], const f32x2 (&e23)[CPL], const float (&s_keep)[NR], float &ent, int dbg)
{
    const float eps = EPS ? TMVB_EPS : 0.0f;
    const f32x2 eps2 = pk2(eps, eps);
#pragma unroll
    for (int j = 0; j < NR; j++) {
        if (rd.c[j] > 0.0f) {
            const float s = s_keep[j];
            const float t = EPS ? fast_div_pos(rd.c[j], s) : __fdividef(rd.c[j], s);
            const f32x2 t2 = pk2(t, t);
            float *srow = stats + (size_t)rd.term[j] * K_ld + 4 * kl;
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                if (4 * (kl + LPT * m) < K) {
                    float px, py, pz, pw;
                    unpk2(mul2(t2, fma2(rd.b[j][m].x, e01[m], eps2)), px, py);
                    unpk2(mul2(t2, fma2(rd.b[j][m].y, e23[m], eps2)), pz, pw);
                    if (!(dbg & 1)) red_add_v4(srow + 4 * LPT * m, px, py, pz, pw);
                }
            }
            if (ELBO && kl == 0) ent += rd.c[j] * __logf(s);
        }
    }
}

// the tile part of a sweep: rounds q = q0, q0 + qstep, ... of the shared-memory tile (n_tile tokens), packed accumulators
template <int LPT, int CPL, bool EPS>
__device__ __forceinline__ void tile_sweep_pk(const float *tile, const float *cnt_s, int n_tile, int RS, f32x2 keps_init, int CH, int ts, int kl, int q0,
                                              int qstep, const f32x2 (&e01)[CPL], const f32x2 (&e23)[CPL], f32x2 (&g01)[CPL], f32x2 (&g23)[CPL],
                                              float &tsum)
{
    static_assert(EPS, "without the epsilon a padded token would divide by zero");
    constexpr int S = 32 / LPT;
    const ulonglong2 zero = make_ulonglong2(0ull, 0ull);
    const int rounds = (n_tile + S - 1) / S;
#pragma unroll 1
    for (int q = q0; q < rounds; q += qstep) {
        const int n = q * S + ts;
        const bool ok = n < n_tile;
        const ulonglong2 *row = reinterpret_cast<const ulonglong2 *>(tile + (ok ? n : 0) * RS) + kl;
        ulonglong2 b[CPL];
#pragma unroll
        for (int m = 0; m < CPL; m++) b[m] = (m < CPL - 1 || kl + LPT * m < CH) ? row[LPT * m] : zero;
        const float c = ok ? cnt_s[n] : 0.0f;
        const float s = tok_dot2<LPT, CPL>(b, e01, e23, keps_init);
        const float t = c * rcp_ftz(s);
        const f32x2 t2 = pk2(t, t);
#pragma unroll
        for (int m = 0; m < CPL; m++) {
            g01[m] = fma2(b[m].x, t2, g01[m]);
            g23[m] = fma2(b[m].y, t2, g23[m]);
        }
        tsum += t;
    }
}

// the tile part of the scatter pass (the "ELBO = 2" form of tok_final: only sum_n c_n ln s_n is accumulated)
template <int LPT, int CPL, bool EPS, bool ELBO>
__device__ __forceinline__ void tile_final_pk(const float *tile, const float *cnt_s, const int *term_s, float *__restrict__ stats, int n_tile,
                                              int RS, int K, int K_ld, int CH, int ts, int kl, int q0, int qstep, const f32x2 (&e01)[CPL],
                                              const f32x2 (&e23)[CPL], float &ent, int dbg)
{
    constexpr int S = 32 / LPT;
    const float Keps = EPS ? (float)K * TMVB_EPS : 0.0f;
    const float eps = EPS ? TMVB_EPS : 0.0f;
    const f32x2 eps2 = pk2(eps, eps);
    const ulonglong2 zero = make_ulonglong2(0ull, 0ull);
    const int rounds = (n_tile + S - 1) / S;
    for (int q = q0; q < rounds; q += qstep) {
        const int n = q * S + ts;
        const bool ok = n < n_tile;
        const ulonglong2 *row = reinterpret_cast<const ulonglong2 *>(tile + (ok ? n : 0) * RS) + kl;
        ulonglong2 b[CPL];
#pragma unroll
        for (int m = 0; m < CPL; m++) b[m] = (m < CPL - 1 || kl + LPT * m < CH) ? row[LPT * m] : zero;
        const float s = tok_dot<LPT, CPL>(b, e01, e23) + Keps;
        if (ok) {
            const float c = cnt_s[n];
            const float t = EPS ? fast_div_pos(c, s) : __fdividef(c, s);
            const f32x2 t2 = pk2(t, t);
            float *srow = stats + (size_t)term_s[n] * K_ld + 4 * kl;
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                if (4 * (kl + LPT * m) < K) {
                    float px, py, pz, pw;
                    unpk2(mul2(t2, fma2(b[m].x, e01[m], eps2)), px, py);
                    unpk2(mul2(t2, fma2(b[m].y, e23[m], eps2)), pz, pw);
                    if (!(dbg & 1)) red_add_v4(srow + 4 * LPT * m, px, py, pz, pw);
                }
            }
            if (ELBO && kl == 0) ent += c * __logf(s);
        }
    }
}

// psi(x) for two arguments at once in packed fp32 (FFMA2 / FMUL2 / FADD2; the MUFU ops stay scalar): the same
// shift-by-6 rational recurrence and 3-term asymptotic series as psi_lgamma<false, true>.
__device__ __forceinline__ void psi_pair(float x0, float x1, float &p0, float &p1)
{
    const f32x2 x = pk2(x0, x1);
    f32x2 P = x, D = pk2(1.0f, 1.0f);
#pragma unroll
    for (int k = 1; k < 6; k++) {
        const f32x2 f = add2(x, pk2((float)k, (float)k));
        D = fma2(D, f, P);
        P = mul2(P, f);
    }
    float P0, P1, D0, D1;
    unpk2(P, P0, P1);
    unpk2(D, D0, D1);
    const bool lo0 = x0 < 6.0f, lo1 = x1 < 6.0f;
    const float y0 = lo0 ? x0 + 6.0f : x0, y1 = lo1 ? x1 + 6.0f : x1;
    const float c0 = lo0 ? __fdividef(D0, P0) : 0.0f, c1 = lo1 ? __fdividef(D1, P1) : 0.0f;
    const f32x2 t = pk2(__fdividef(1.0f, y0), __fdividef(1.0f, y1)), nt2 = mul2(mul2(t, t), pk2(-1.0f, -1.0f));
    // psi = ln y - t/2 - t2 (1/12 - t2 (1/120 - t2/252)) - corr
    const f32x2 in1 = fma2(nt2, pk2(3.9682539683e-3f, 3.9682539683e-3f), pk2(8.3333333333e-3f, 8.3333333333e-3f));
    const f32x2 in2 = fma2(nt2, in1, pk2(8.3333333333e-2f, 8.3333333333e-2f));
    f32x2 r = fma2(t, pk2(-0.5f, -0.5f), pk2(__logf(y0) - c0, __logf(y1) - c1));
    r = fma2(nt2, in2, r);
    unpk2(r, p0, p1);
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

// owner-lane sums of two consecutive topics i, i+1 (i even): LDS.64 + FADD2, four independent chains
template <int S>
__device__ __forceinline__ float2 owner_sum2(const float *gs, int RS, int i)
{
    f32x2 g0 = 0ull, g1 = 0ull, g2 = 0ull, g3 = 0ull;
    if (S >= 4) {
#pragma unroll
        for (int w = 0; w < S; w += 4) {
            g0 = add2(g0, *reinterpret_cast<const f32x2 *>(gs + w * RS + i));
            g1 = add2(g1, *reinterpret_cast<const f32x2 *>(gs + (w + 1) * RS + i));
            g2 = add2(g2, *reinterpret_cast<const f32x2 *>(gs + (w + 2) * RS + i));
            g3 = add2(g3, *reinterpret_cast<const f32x2 *>(gs + (w + 3) * RS + i));
        }
    } else {
#pragma unroll
        for (int w = 0; w < S; w++) g0 = add2(g0, *reinterpret_cast<const f32x2 *>(gs + w * RS + i));
    }
    float2 r;
    unpk2(add2(add2(g0, g1), add2(g2, g3)), r.x, r.y);
    return r;
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
