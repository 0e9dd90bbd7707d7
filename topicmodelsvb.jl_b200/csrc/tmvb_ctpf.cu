// tmvb_ctpf.cu -- collaborative topic Poisson factorization coordinate-ascent VB on sm_100a and the tmvb_ctpf_* C ABI.
//
// Reference semantics: the CPU model src/CTPF.jl (per-document stopping rule ||gimel - gimel_old|| < vtol, `vav` in
// update_xi where the OpenCL kernel has `bet` (gpuCTPF.jl:624 vs CTPF.jl:336), lagged phi/xi ELBO).  Replaced:
// src/gpuCTPF.jl's 12 OpenCL kernels + modelutils.jl:438-494,540-570.
//
// The two softmaxes of a sweep share the structure of tmvb_estep.cuh:
//   phi_n  = softmax_i(psi(gimel_i) - ln dalet_i - ln bet_i + psi(alef[i,w_n]))            (CTPF.jl:327-330)
//          = A[i,w_n] ephi_i / s_n,   A = exp(psi(alef) - rowmax)  (a K x V table rebuilt once per outer iteration
//            instead of one digamma per (topic, token, sweep), gpuCTPF.jl:560),  ephi = exp(psi(gimel) - ln dalet - ln bet - max)
//   xi_r   = softmax over 2K of [psi(gimel) - ln dalet - ln vav + psi(he[:,r]) ; psi(zayin) - ln het - ln vav + psi(he[:,r])]
//          = [H[:,r] ea ; H[:,r] eb] / s_r,   H = exp(psi(he) - rowmax) (K x U table)              (CTPF.jl:334-337)
// so gimel = c + ephi .* gphi + ea .* gx and zayin = g + eb .* gx with gphi = sum_n A[:,w_n] c_n/s_n, gx = sum_r H[:,r] rating_r/s_r
// (CTPF.jl:309-323): two token passes per sweep over two shared-memory tiles (term rows of A, reader rows of H).
// Scatters (CTPF.jl:259-262,274-277) and rate updates (CTPF.jl:281-305) follow; the global Gamma-rate algebra runs in fp64
// on the host from K-vectors.
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TMVB_RECS_KERNELS
#include "tmvb_comm.cuh"
#include "tmvb_recs.cuh"
#include "tmvb_shard.cuh"

namespace tmvb {

enum { KV_LNDB = 0, KV_LNDV, KV_LNHV, KV_LN_DALET, KV_LN_BET, KV_LN_VAV, KV_LN_HET, KV_AL_DB, KV_HE_DV, KV_HE_HV, KV_INV_DALET, KV_INV_HET,
       KV_LNDB_OLD, KV_LNDV_OLD, KV_LNHV_OLD, KV_COUNT };

struct CtpfDev {
    int K, K_ld, V, U, RS;
    long long M;
    const float *A;       // [V][K_ld] exp(psi(alef) - rowmax)
    const float *H;       // [U][K_ld] exp(psi(he) - rowmax)
    float *stats_a;       // [V][K_ld]
    float *stats_h;       // [U][K_ld]
    const long long *doc_off, *r_off;
    const int *terms, *readers;
    const float *counts, *ratings;
    float *gimel, *gimel_old, *zayin, *zayin_old;  // [M][K_ld]
    const float *kv;      // [KV_COUNT][K_ld]
    double *small;        // [K_ld] sum gimel | [K_ld] sum zayin | [1] per-document ELBO terms | [1] sweeps
    float hc, hg;         // hyper-parameters c and g (shape priors of theta and epsilon)
    int viter;
    float vtol;
    int stage_bulk, dbg;
};

static size_t ctpf_fixed_smem(int RS, int lpt) { return 16 + (size_t)(32 / lpt) * RS * 4 + (size_t)RS * 4; }

__device__ __forceinline__ float warp_max_f(float v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

template <int LPT, int CPL, bool ELBO>
__global__ void __launch_bounds__(32) ctpf_estep_kernel(const CtpfDev p, int doc_begin, int doc_end, int cap, int cap2, int *counter)
{
    constexpr int S = 32 / LPT;
    constexpr int R = (LPT * CPL + 7) / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int kl = lane % LPT, ts = lane / LPT;
    const int K = p.K, K_ld = p.K_ld, CH = K_ld >> 2, RS = p.RS;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw);
    float *gs = reinterpret_cast<float *>(smem_raw + 16);  // [S][RS]
    float *e_s = gs + (size_t)S * RS;                      // [RS]
    float *tile = e_s + RS;                                // [cap][RS]   term rows of A
    float *tile2 = tile + (size_t)cap * RS;                // [cap2][RS]  reader rows of H
    float *cnt_s = tile2 + (size_t)cap2 * RS;              // [cap]
    float *cnt2_s = cnt_s + cap;                           // [cap2]
    int *term_s = reinterpret_cast<int *>(cnt2_s + cap2);  // [cap]
    int *term2_s = term_s + cap;                           // [cap2]

    float lndb_k[R], lndv_k[R], lnhv_k[R];
    double gs_k[R], zs_k[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = lane + 32 * r;
        lndb_k[r] = (i < K) ? p.kv[KV_LNDB * K_ld + i] : 0.0f;
        lndv_k[r] = (i < K) ? p.kv[KV_LNDV * K_ld + i] : 0.0f;
        lnhv_k[r] = (i < K) ? p.kv[KV_LNHV * K_ld + i] : 0.0f;
        gs_k[r] = zs_k[r] = 0.0;
    }
    const float dscale = (p.vtol > 0.0f) ? 1048576.0f / (p.vtol * p.vtol) : 0.0f;
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
        const long long o = p.doc_off[d], ro = p.r_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o), Rd = (int)(p.r_off[d + 1] - ro);
        const int ns = min(Nd, cap), nr = min(Rd, cap2);
        const bool ovf = Nd > cap, ovf2 = Rd > cap2;

        // stage both tiles under one mbarrier transaction
        __syncwarp();
        float lg_c = 0.0f;  // sum_n lnG(c_n + 1) + sum_r lnG(rating_r + 1)
        for (int n = lane; n < Nd; n += 32) {
            const float c = p.counts[o + n];
            if (ELBO && c > 1.5f) lg_c += lgammaf(c + 1.0f);
            if (n < ns) {
                term_s[n] = p.terms[o + n];
                cnt_s[n] = c;
            }
        }
        for (int n = lane; n < Rd; n += 32) {
            const float c = p.ratings[ro + n];
            if (ELBO && c > 1.5f) lg_c += lgammaf(c + 1.0f);
            if (n < nr) {
                term2_s[n] = p.readers[ro + n];
                cnt2_s[n] = c;
            }
        }
        if (p.stage_bulk) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(mbar, (unsigned)((ns + nr) * K_ld * 4));
            __syncwarp();
            for (int n = lane; n < ns; n += 32) bulk_g2s(tile + n * RS, p.A + (size_t)term_s[n] * K_ld, (unsigned)(K_ld * 4), mbar);
            for (int n = lane; n < nr; n += 32) bulk_g2s(tile2 + n * RS, p.H + (size_t)term2_s[n] * K_ld, (unsigned)(K_ld * 4), mbar);
        } else {
            __syncwarp();
            for (int c = lane; c < ns * CH; c += 32) {
                const int n = c / CH, q = c - n * CH;
                cp_async16(tile + n * RS + 4 * q, p.A + (size_t)term_s[n] * K_ld + 4 * q);
            }
            for (int c = lane; c < nr * CH; c += 32) {
                const int n = c / CH, q = c - n * CH;
                cp_async16(tile2 + n * RS + 4 * q, p.H + (size_t)term2_s[n] * K_ld + 4 * q);
            }
            cp_async_commit();
        }
        float gim_k[R], gimo_k[R], zay_k[R], zayo_k[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            gim_k[r] = (i < K) ? p.gimel[(size_t)d * K_ld + i] : 1.0f;
            zay_k[r] = (i < K) ? p.zayin[(size_t)d * K_ld + i] : 1.0f;
            gimo_k[r] = gim_k[r];
            zayo_k[r] = zay_k[r];
        }
        stage_wait(mbar, phase, p.stage_bulk);

        TokArgs ta, tb;
        ta.tile = tile;
        ta.cnt_s = cnt_s;
        ta.term_s = term_s;
        ta.gtable = p.A;
        ta.gterms = p.terms + o;
        ta.gcounts = p.counts + o;
        ta.stats = p.stats_a;
        ta.Nd = Nd;
        ta.cap = cap;
        ta.rounds = (Nd + S - 1) / S;
        ta.K = K;
        ta.K_ld = K_ld;
        ta.RS = RS;
        ta.dbg = p.dbg;
        ta.r0 = 0;
        ta.rstep = 1;
        tb = ta;
        tb.tile = tile2;
        tb.cnt_s = cnt2_s;
        tb.term_s = term2_s;
        tb.gtable = p.H;
        tb.gterms = p.readers + ro;
        tb.gcounts = p.ratings + ro;
        tb.stats = p.stats_h;
        tb.Nd = Rd;
        tb.cap = cap2;
        tb.rounds = (Rd + S - 1) / S;

        float ephi_k[R], ea_k[R], eb_k[R];
        int v = 0;
        for (;;) {
            // ---- the per-document halves of update_xi! / update_phi! (CTPF.jl:327-337)
            float mphi = -INFINITY, mx = -INFINITY;
            float aphi[R], aa[R], ab[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const float pg = psi_lgamma<false, true>(gim_k[r]).psi, pz = psi_lgamma<false, true>(zay_k[r]).psi;
                aphi[r] = pg + lndb_k[r];
                aa[r] = pg + lndv_k[r];
                ab[r] = pz + lnhv_k[r];
                if (lane + 32 * r < K) {
                    mphi = fmaxf(mphi, aphi[r]);
                    mx = fmaxf(mx, fmaxf(aa[r], ab[r]));
                }
            }
            mphi = warp_max_f(mphi);
            mx = warp_max_f(mx);
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = lane + 32 * r;
                ephi_k[r] = (i < K) ? __expf(aphi[r] - mphi) : 0.0f;
                ea_k[r] = (i < K) ? __expf(aa[r] - mx) : 0.0f;
                eb_k[r] = (i < K) ? __expf(ab[r] - mx) : 0.0f;
                if (i < K_ld) e_s[i] = ephi_k[r];
            }
            __syncwarp();
            float4 e[CPL], g[CPL];
            float tsum = 0.0f;
            // ---- phi pass over the term tile
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                e[m] = (m < CPL - 1 || kl + LPT * m < CH) ? reinterpret_cast<const float4 *>(e_s)[kl + LPT * m] : zero4;
                g[m] = zero4;
            }
            if (!ovf)
                tok_sweep<LPT, CPL, false, false>(ta, ts, kl, e, g, tsum);
            else
                tok_sweep<LPT, CPL, true, false>(ta, ts, kl, e, g, tsum);
#pragma unroll
            for (int m = 0; m < CPL; m++)
                if (m < CPL - 1 || kl + LPT * m < CH) reinterpret_cast<float4 *>(gs + ts * RS)[kl + LPT * m] = g[m];
            __syncwarp();
            float gphi[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = lane + 32 * r;
                gphi[r] = (i < K) ? owner_sum<S>(gs, RS, i) : 0.0f;
                if (i < K_ld) e_s[i] = ea_k[r] + eb_k[r];
            }
            __syncwarp();
            // ---- xi pass over the reader tile
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                e[m] = (m < CPL - 1 || kl + LPT * m < CH) ? reinterpret_cast<const float4 *>(e_s)[kl + LPT * m] : zero4;
                g[m] = zero4;
            }
            if (!ovf2)
                tok_sweep<LPT, CPL, false, false>(tb, ts, kl, e, g, tsum);
            else
                tok_sweep<LPT, CPL, true, false>(tb, ts, kl, e, g, tsum);
#pragma unroll
            for (int m = 0; m < CPL; m++)
                if (m < CPL - 1 || kl + LPT * m < CH) reinterpret_cast<float4 *>(gs + ts * RS)[kl + LPT * m] = g[m];
            __syncwarp();
            // ---- update_zayin!, update_gimel! (CTPF.jl:309-323)
            float dpart = 0.0f;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = lane + 32 * r;
                const float gx = (i < K) ? owner_sum<S>(gs, RS, i) : 0.0f;
                zayo_k[r] = zay_k[r];
                gimo_k[r] = gim_k[r];
                if (i < K) {
                    zay_k[r] = p.hg + eb_k[r] * gx;
                    gim_k[r] = p.hc + fmaf(ephi_k[r], gphi[r], ea_k[r] * gx);
                    const float df = gim_k[r] - gimo_k[r];
                    dpart = fmaf(df, df, dpart);
                }
            }
            __syncwarp();
            v++;
            if (v >= p.viter) break;  // CTPF.jl:359
            if (dscale > 0.0f && __reduce_add_sync(0xffffffffu, (unsigned)fminf(dpart * dscale, 67108864.0f)) < 1048576u) break;
        }

        // ---- update_alef!(d) and update_he!(d) (CTPF.jl:259-262,274-277): scatter the last phi / xi
        float ent = 0.0f;
        {
            float4 e[CPL], e2[CPL];
#pragma unroll
            for (int r = 0; r < R; r++)
                if (lane + 32 * r < K_ld) e_s[lane + 32 * r] = ephi_k[r];
            __syncwarp();
#pragma unroll
            for (int m = 0; m < CPL; m++)
                e[m] = (m < CPL - 1 || kl + LPT * m < CH) ? reinterpret_cast<const float4 *>(e_s)[kl + LPT * m] : zero4;
            if (!ovf)
                tok_final<LPT, CPL, false, false, ELBO>(ta, ts, kl, e, ent);
            else
                tok_final<LPT, CPL, true, false, ELBO>(ta, ts, kl, e, ent);
            if (Rd > 0) {
                __syncwarp();
#pragma unroll
                for (int r = 0; r < R; r++)
                    if (lane + 32 * r < K_ld) e_s[lane + 32 * r] = ea_k[r];
                __syncwarp();
#pragma unroll
                for (int m = 0; m < CPL; m++)
                    e[m] = (m < CPL - 1 || kl + LPT * m < CH) ? reinterpret_cast<const float4 *>(e_s)[kl + LPT * m] : zero4;
                __syncwarp();
#pragma unroll
                for (int r = 0; r < R; r++)
                    if (lane + 32 * r < K_ld) e_s[lane + 32 * r] = eb_k[r];
                __syncwarp();
#pragma unroll
                for (int m = 0; m < CPL; m++)
                    e2[m] = (m < CPL - 1 || kl + LPT * m < CH) ? reinterpret_cast<const float4 *>(e_s)[kl + LPT * m] : zero4;
                if (!ovf2)
                    tok_final2<LPT, CPL, false, ELBO>(tb, ts, kl, e, e2, ent);
                else
                    tok_final2<LPT, CPL, true, ELBO>(tb, ts, kl, e, e2, ent);
            }
        }

        float a = 0.0f;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            if (i < K_ld) {
                const bool ok = i < K;
                p.gimel[(size_t)d * K_ld + i] = ok ? gim_k[r] : 0.0f;
                p.gimel_old[(size_t)d * K_ld + i] = ok ? gimo_k[r] : 0.0f;
                p.zayin[(size_t)d * K_ld + i] = ok ? zay_k[r] : 0.0f;
                p.zayin_old[(size_t)d * K_ld + i] = ok ? zayo_k[r] : 0.0f;
                if (ok) {
                    gs_k[r] += (double)gim_k[r];
                    zs_k[r] += (double)zay_k[r];
                    // every psi(gimel), psi(zayin) term of the ELBO cancels (see tmvb_ctpf_elbo); what is left per
                    // document is sum_i lnG(gimel_i) + lnG(zayin_i) and the z / y entropies
                    if (ELBO) a += psi_lgamma<true>(gim_k[r]).lg + psi_lgamma<true>(zay_k[r]).lg;
                }
            }
        }
        if (ELBO) elbo_thr += (double)a + (double)ent - (double)lg_c;
        if (lane == 0) sweeps_thr += (unsigned long long)v;
    }

    if (ELBO) {
        const double tot = warp_sum_d(elbo_thr);
        if (lane == 0 && tot != 0.0) atomicAdd(p.small + 2 * K_ld, tot);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = lane + 32 * r;
        if (i < K) {
            if (gs_k[r] != 0.0) atomicAdd(p.small + i, gs_k[r]);
            if (zs_k[r] != 0.0) atomicAdd(p.small + K_ld + i, zs_k[r]);
        }
    }
    if (lane == 0 && sweeps_thr) atomicAdd(p.small + 2 * K_ld + 1, (double)sweeps_thr);
}

// x = prior + stats (one warp per table row): raw <- x, table <- exp(psi(x) - rowmax), stats <- 0, and per-topic sums
// acc = [sum psi(x) (K_ld) | sum x (K_ld) | sum lnG(x) + (1-x) psi(x) (K_ld) | sum stats psi(x) (1)]
// (update_alef!/update_he! CTPF.jl:251-270 + the table of exp(psi) the sweeps read + the global ELBO sums CTPF.jl:143-168,197-221)
// RMAX = topics per lane (ceil(K / 32) rounded up to 1 / 2 / 4 / 8).  The per-topic sums are reduced in shared memory per CTA before
// they reach the global accumulators (one warp per row with its own 3 K global fp64 atomics made the 8 000-row alef table cost 65 us).
template <int RMAX>
__global__ void ctpf_table_kernel(float *__restrict__ stats, float prior, float *__restrict__ raw, float *__restrict__ table, int rows,
                                  int K, int K_ld, double *__restrict__ acc, int zero_stats)
{
    extern __shared__ double tsh[];   // [3 K_ld + 1]
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int q = threadIdx.x; q < 3 * K_ld + 1; q += blockDim.x) tsh[q] = 0.0;
    __syncthreads();
    double t1[RMAX], t2[RMAX], t3[RMAX], t4 = 0.0;
#pragma unroll
    for (int r = 0; r < RMAX; r++) t1[r] = t2[r] = t3[r] = 0.0;
    for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < rows; j += gridDim.x * wpb) {
        float ps[RMAX], x[RMAX];
        float mx = -INFINITY;
#pragma unroll
        for (int r = 0; r < RMAX; r++) {
            const int i = lane + 32 * r;
            ps[r] = 0.0f;
            x[r] = 0.0f;
            if (i < K) {
                const float sv = stats[(size_t)j * K_ld + i];
                x[r] = prior + sv;
                const PsiLg pl = psi_lgamma<true>(x[r]);
                ps[r] = pl.psi;
                mx = fmaxf(mx, ps[r]);
                t1[r] += (double)ps[r];
                t2[r] += (double)x[r];
                t3[r] += (double)(pl.lg + (1.0f - x[r]) * ps[r]);
                t4 += (double)(sv * ps[r]);
            }
        }
        mx = warp_max_f(mx);
#pragma unroll
        for (int r = 0; r < RMAX; r++) {
            const int i = lane + 32 * r;
            if (i < K_ld) {
                raw[(size_t)j * K_ld + i] = (i < K) ? x[r] : 0.0f;
                table[(size_t)j * K_ld + i] = (i < K) ? expf(ps[r] - mx) : 0.0f;
                if (zero_stats) stats[(size_t)j * K_ld + i] = 0.0f;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RMAX; r++) {
        const int i = lane + 32 * r;
        if (i < K) {
            atomicAdd(tsh + i, t1[r]);
            atomicAdd(tsh + K_ld + i, t2[r]);
            atomicAdd(tsh + 2 * K_ld + i, t3[r]);
        }
    }
    t4 = warp_sum_d(t4);
    if (lane == 0 && t4 != 0.0) atomicAdd(tsh + 3 * K_ld, t4);
    __syncthreads();
    for (int q = threadIdx.x; q < 3 * K_ld + 1; q += blockDim.x)
        if (tsh[q] != 0.0) atomicAdd(acc + q, tsh[q]);
}

// update_elbo! restated (CTPF.jl:232-247) per document: phi / xi from the *_old state (A_old, H_old, gimel_old, zayin_old,
// old rates), every expectation with the current alef / he / rates / gimel / zayin.  One warp per document, lanes over
// topics; fp32 per element, fp64 accumulation.  The corpus-level Gamma terms of alef / he are added on the host.
// Everything that depends on the document only -- exp(psi(gimel_old) + ln rates_old - max) of the three softmaxes and the
// psi(gimel) / psi(zayin) - ln rates constants of the expectations -- is evaluated once per document and kept in RM registers per
// lane (topics i = lane + 32 r); per (token, topic) and (reader, topic) there remain one table load, psi(alef) resp. psi(he) and
// one logarithm (the first version re-evaluated four digammas and two exponentials per element: 2.1 ms at CiteULike, three
// E-steps' worth, inside every train! call).
template <int RM>
__global__ void ctpf_elbo_kernel(const CtpfDev p, const float *__restrict__ A_old, const float *__restrict__ H_old,
                                 const float *__restrict__ alef, const float *__restrict__ he, float hd, float hh, double *out)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int K = p.K, K_ld = p.K_ld;
    const float *kv = p.kv;
    double acc = 0.0;
    for (long long d = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); d < p.M; d += (long long)gridDim.x * wpb) {
        const long long o = p.doc_off[d], ro = p.r_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o), Rd = (int)(p.r_off[d + 1] - ro);
        const float *gm = p.gimel + d * K_ld, *zy = p.zayin + d * K_ld, *go = p.gimel_old + d * K_ld, *zo = p.zayin_old + d * K_ld;
        float xphi[RM], xa[RM], xb[RM], cphi[RM], ca[RM], cb[RM];
        float mphi = -INFINITY, mx = -INFINITY, b = 0.0f;
#pragma unroll
        for (int r = 0; r < RM; r++) {
            const int i = lane + 32 * r;
            xphi[r] = xa[r] = xb[r] = -INFINITY;
            cphi[r] = ca[r] = cb[r] = 0.0f;
            if (i < K) {
                const float pgo = psi_lgamma<false>(go[i]).psi, pzo = psi_lgamma<false>(zo[i]).psi;
                xphi[r] = pgo + kv[KV_LNDB_OLD * K_ld + i];
                xa[r] = pgo + kv[KV_LNDV_OLD * K_ld + i];
                xb[r] = pzo + kv[KV_LNHV_OLD * K_ld + i];
                mphi = fmaxf(mphi, xphi[r]);
                mx = fmaxf(mx, fmaxf(xa[r], xb[r]));
                const PsiLg pg = psi_lgamma<true>(gm[i]), pz = psi_lgamma<true>(zy[i]);
                cphi[r] = pg.psi - kv[KV_LN_DALET * K_ld + i] - kv[KV_LN_BET * K_ld + i];
                ca[r] = pg.psi - kv[KV_LN_DALET * K_ld + i] - kv[KV_LN_VAV * K_ld + i];
                cb[r] = pz.psi - kv[KV_LN_HET * K_ld + i] - kv[KV_LN_VAV * K_ld + i];
                b -= gm[i] * (kv[KV_HE_DV * K_ld + i] + kv[KV_AL_DB * K_ld + i]) + zy[i] * kv[KV_HE_HV * K_ld + i];   // CTPF.jl:112,123,134
                b += (p.hc - 1.0f) * (pg.psi - kv[KV_LN_DALET * K_ld + i]) - hd * gm[i] * kv[KV_INV_DALET * K_ld + i];  // Elogptheta
                b += (p.hg - 1.0f) * (pz.psi - kv[KV_LN_HET * K_ld + i]) - hh * zy[i] * kv[KV_INV_HET * K_ld + i];      // Elogpepsilon
                b += gm[i] - kv[KV_LN_DALET * K_ld + i] + pg.lg + (1.0f - gm[i]) * pg.psi;                              // entropy(Gamma)
                b += zy[i] - kv[KV_LN_HET * K_ld + i] + pz.lg + (1.0f - zy[i]) * pz.psi;
            }
        }
        mphi = warp_max_f(mphi);
        mx = warp_max_f(mx);
#pragma unroll
        for (int r = 0; r < RM; r++) {   // exp(-inf) = 0 for the slots beyond K
            xphi[r] = expf(xphi[r] - mphi);
            xa[r] = expf(xa[r] - mx);
            xb[r] = expf(xb[r] - mx);
        }
        double dacc = 0.0;
        for (int n = 0; n < Nd; n++) {
            const int term = p.terms[o + n];
            const float c = p.counts[o + n];
            const float *Ao = A_old + (size_t)term * K_ld, *al = alef + (size_t)term * K_ld;
            float u[RM], s = 0.0f;
#pragma unroll
            for (int r = 0; r < RM; r++) {
                const int i = lane + 32 * r;
                u[r] = (i < K) ? Ao[i] * xphi[r] : 0.0f;
                s += u[r];
            }
            s = warp_sum(s);
            const float rs = 1.0f / s;
            float a = 0.0f;
#pragma unroll
            for (int r = 0; r < RM; r++) {
                const int i = lane + 32 * r;
                const float ph = u[r] * rs;
                if (i < K && ph > 0.0f) a += ph * (cphi[r] + psi_lgamma<false>(al[i]).psi - logf(ph));
            }
            dacc += (double)(c * a);
            if (lane == 0) dacc -= (double)lgammaf(c + 1.0f);
        }
        for (int n = 0; n < Rd; n++) {
            const int uu = p.readers[ro + n];
            const float c = p.ratings[ro + n];
            const float *Ho = H_old + (size_t)uu * K_ld, *hr = he + (size_t)uu * K_ld;
            float va[RM], vb[RM], s = 0.0f;
#pragma unroll
            for (int r = 0; r < RM; r++) {
                const int i = lane + 32 * r;
                const float hv = (i < K) ? Ho[i] : 0.0f;
                va[r] = hv * xa[r];
                vb[r] = hv * xb[r];
                s += va[r] + vb[r];
            }
            s = warp_sum(s);
            const float rs = 1.0f / s;
            float a = 0.0f;
#pragma unroll
            for (int r = 0; r < RM; r++) {
                const int i = lane + 32 * r;
                if (i < K) {
                    const float phe = psi_lgamma<false>(hr[i]).psi, pa = va[r] * rs, pb = vb[r] * rs;
                    if (pa > 0.0f) a += pa * (ca[r] + phe - logf(pa));
                    if (pb > 0.0f) a += pb * (cb[r] + phe - logf(pb));
                }
            }
            dacc += (double)(c * a);
            if (lane == 0) dacc -= (double)lgammaf(c + 1.0f);
        }
        dacc += (double)b;
        dacc = warp_sum_d(dacc);
        if (lane == 0) acc += dacc;
    }
    if (lane == 0 && acc != 0.0) atomicAdd(out, acc);
}

typedef void (*CtpfEstepFn)(const CtpfDev, int, int, int, int, int *);
#define TMVB_CTPF_FN(L, C) {(CtpfEstepFn)ctpf_estep_kernel<L, C, false>, (CtpfEstepFn)ctpf_estep_kernel<L, C, true>},
static const CtpfEstepFn kCtpfEstep[kNumLaneLayouts][2] = {TMVB_FOR_EACH_LAYOUT(TMVB_CTPF_FN)};

}  // namespace tmvb

using namespace tmvb;

struct tmvb_ctpf_s {
    Shard s;  // d_beta[2] = A tables (current / previous), d_stats = alef statistics
    int64_t U = 0, nnz_r = 0;
    bool elbo_valid = false, readers_set = false;
    double hyp[8] = {0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1};  // a b c d e f g h (gpuCTPF.jl:107)
    float *d_alef = nullptr, *d_alef_old = nullptr;            // raw [V][K_ld]
    float *d_he = nullptr, *d_he_old = nullptr;                // raw [U][K_ld]
    float *d_H[2] = {nullptr, nullptr}, *d_hstats = nullptr;   // [U][K_ld]
    int hcur = 0;
    long long *d_r_off = nullptr;
    int *d_readers = nullptr;
    float *d_ratings = nullptr;
    float *d_gimel = nullptr, *d_gimel_old = nullptr, *d_zayin = nullptr, *d_zayin_old = nullptr;
    float *d_kv = nullptr;
    std::vector<double> bet, vav, dalet, het, bet_old, vav_old, dalet_old, het_old;  // fp64 masters
    double *d_small = nullptr;  // [2 K_ld + 2]
    double *d_tsum = nullptr;   // [2][3 K_ld + 1] table sums (alef | he) + [1] standalone ELBO
    std::vector<double> h_small, h_tsum;
    std::vector<int> r_len;
    Comm comm;   // peer-memory all-reduce of the statistics (multi-GPU)
};

namespace {

CtpfDev ctpf_view(tmvb_ctpf_t h)
{
    Shard &s = h->s;
    CtpfDev p;
    memset(&p, 0, sizeof(p));  // the struct is also the key of the captured launch graph: no indeterminate padding
    p.K = (int)s.K;
    p.K_ld = s.K_ld;
    p.V = (int)s.V;
    p.U = (int)h->U;
    p.RS = s.RS;
    p.M = s.M;
    p.A = s.d_beta[s.cur];
    p.H = h->d_H[h->hcur];
    p.stats_a = s.d_stats;
    p.stats_h = h->d_hstats;
    p.doc_off = s.d_doc_off;
    p.r_off = h->d_r_off;
    p.terms = s.d_terms;
    p.readers = h->d_readers;
    p.counts = s.d_counts;
    p.ratings = h->d_ratings;
    p.gimel = h->d_gimel;
    p.gimel_old = h->d_gimel_old;
    p.zayin = h->d_zayin;
    p.zayin_old = h->d_zayin_old;
    p.kv = h->d_kv;
    p.small = h->d_small;
    p.hc = (float)h->hyp[2];
    p.hg = (float)h->hyp[6];
    p.viter = 0;
    p.vtol = 0.f;
    p.stage_bulk = env_int("TMVB_STAGE_BULK", 1);
    p.dbg = env_int("TMVB_DBG", 0);
    return p;
}

void ctpf_free(tmvb_ctpf_t h)
{
    cudaSetDevice(h->s.device);
    if (h->s.stream) cudaStreamSynchronize(h->s.stream);
    comm_free(&h->comm);
    cudaFree(h->d_alef);
    cudaFree(h->d_alef_old);
    cudaFree(h->d_he);
    cudaFree(h->d_he_old);
    cudaFree(h->d_H[0]);
    cudaFree(h->d_H[1]);
    cudaFree(h->d_hstats);
    cudaFree(h->d_r_off);
    cudaFree(h->d_readers);
    cudaFree(h->d_ratings);
    cudaFree(h->d_gimel);
    cudaFree(h->d_gimel_old);
    cudaFree(h->d_zayin);
    cudaFree(h->d_zayin_old);
    cudaFree(h->d_kv);
    cudaFree(h->d_small);
    cudaFree(h->d_tsum);
    shard_free(&h->s);
}

// K-vectors the kernels read, from the fp64 masters; AL / HE are the row sums of the current alef / he
int ctpf_push_kv(tmvb_ctpf_t h)
{
    Shard &s = h->s;
    const int K = (int)s.K, K_ld = s.K_ld;
    std::vector<float> kv((size_t)KV_COUNT * K_ld, 0.f);
    const double *AL = h->h_tsum.data() + K_ld, *HE = h->h_tsum.data() + (3 * K_ld + 1) + K_ld;
    for (int i = 0; i < K; i++) {
        kv[KV_LNDB * K_ld + i] = (float)(-log(h->dalet[i]) - log(h->bet[i]));
        kv[KV_LNDV * K_ld + i] = (float)(-log(h->dalet[i]) - log(h->vav[i]));
        kv[KV_LNHV * K_ld + i] = (float)(-log(h->het[i]) - log(h->vav[i]));
        kv[KV_LN_DALET * K_ld + i] = (float)log(h->dalet[i]);
        kv[KV_LN_BET * K_ld + i] = (float)log(h->bet[i]);
        kv[KV_LN_VAV * K_ld + i] = (float)log(h->vav[i]);
        kv[KV_LN_HET * K_ld + i] = (float)log(h->het[i]);
        kv[KV_AL_DB * K_ld + i] = (float)(AL[i] / (h->dalet[i] * h->bet[i]));
        kv[KV_HE_DV * K_ld + i] = (float)(HE[i] / (h->dalet[i] * h->vav[i]));
        kv[KV_HE_HV * K_ld + i] = (float)(HE[i] / (h->het[i] * h->vav[i]));
        kv[KV_INV_DALET * K_ld + i] = (float)(1.0 / h->dalet[i]);
        kv[KV_INV_HET * K_ld + i] = (float)(1.0 / h->het[i]);
        kv[KV_LNDB_OLD * K_ld + i] = (float)(-log(h->dalet_old[i]) - log(h->bet_old[i]));
        kv[KV_LNDV_OLD * K_ld + i] = (float)(-log(h->dalet_old[i]) - log(h->vav_old[i]));
        kv[KV_LNHV_OLD * K_ld + i] = (float)(-log(h->het_old[i]) - log(h->vav_old[i]));
    }
    TMVB_CUDA(cudaMemcpyAsync(h->d_kv, kv.data(), kv.size() * 4, cudaMemcpyHostToDevice, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    s.st.h2d_bytes += (int64_t)kv.size() * 4;
    return 0;
}

// run the table kernel for alef (which = 0) or he (which = 1); results land in h_tsum after the caller's D2H
int ctpf_build_table(tmvb_ctpf_t h, int which, float prior, int zero_stats)
{
    Shard &s = h->s;
    const int rows = which == 0 ? (int)s.V : (int)h->U;
    double *acc = h->d_tsum + (size_t)which * (3 * s.K_ld + 1);
    TMVB_CUDA(cudaMemsetAsync(acc, 0, (3 * s.K_ld + 1) * 8, s.stream));
    if (rows == 0) return 0;
    float *stats = which == 0 ? s.d_stats : h->d_hstats;
    float *raw = which == 0 ? h->d_alef : h->d_he;
    float *table = which == 0 ? s.d_beta[s.cur] : h->d_H[h->hcur];
    const int grid = std::max(1, std::min((rows + 31) / 32, s.n_sm * 4));   // >= 4 rows per warp
    const size_t tsm = (size_t)(3 * s.K_ld + 1) * 8;
#define TMVB_CTPF_TABLE(R) ctpf_table_kernel<R><<<grid, 256, tsm, s.stream>>>(stats, prior, raw, table, rows, (int)s.K, s.K_ld, acc, zero_stats)
    if (s.K <= 32) TMVB_CTPF_TABLE(1); else if (s.K <= 64) TMVB_CTPF_TABLE(2); else if (s.K <= 128) TMVB_CTPF_TABLE(4); else TMVB_CTPF_TABLE(8);
#undef TMVB_CTPF_TABLE
    TMVB_CUDA(cudaGetLastError());
    s.st.kernel_launches++;
    return 0;
}

int ctpf_fetch_tsum(tmvb_ctpf_t h)
{
    Shard &s = h->s;
    const size_t n = 2 * (3 * s.K_ld + 1);
    TMVB_CUDA(cudaMemcpyAsync(s.h_pinned, h->d_tsum, n * 8, cudaMemcpyDeviceToHost, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    memcpy(h->h_tsum.data(), s.h_pinned, n * 8);
    s.st.d2h_bytes += n * 8;
    return 0;
}

}  // namespace

extern "C" {

int tmvb_ctpf_create(tmvb_ctpf_t *out, int64_t K, int64_t M, int64_t V, int64_t U, int device, void *stream)
{
    TMVB_CHECK_ARG(out != nullptr, "handle pointer is NULL");
    *out = nullptr;
    TMVB_CHECK_ARG(U >= 0 && U < (1ll << 31), "U must be nonnegative and fit in int32");
    tmvb_ctpf_t h = new tmvb_ctpf_s();
    const int64_t K_ld = (K + 7) / 8 * 8;
    int rc = shard_create(&h->s, K, M, V, device, stream, (size_t)(8 * K_ld + 16));
    if (rc == 0) {
        Shard &s = h->s;
        h->U = U;
        const size_t kv = (size_t)std::max<int64_t>(V, 1) * s.K_ld, ku = (size_t)std::max<int64_t>(U, 1) * s.K_ld,
                     km = (size_t)std::max<int64_t>(M, 1) * s.K_ld;
        cudaError_t e = cudaSuccess;
        auto A = [&](void **p, size_t bytes) {
            if (e == cudaSuccess) e = cudaMalloc(p, bytes);
            if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes, s.stream);
        };
        A((void **)&h->d_alef, kv * 4);
        A((void **)&h->d_alef_old, kv * 4);
        A((void **)&h->d_he, ku * 4);
        A((void **)&h->d_he_old, ku * 4);
        A((void **)&h->d_H[0], ku * 4);
        A((void **)&h->d_H[1], ku * 4);
        A((void **)&h->d_hstats, ku * 4);
        A((void **)&h->d_gimel, km * 4);
        A((void **)&h->d_gimel_old, km * 4);
        A((void **)&h->d_zayin, km * 4);
        A((void **)&h->d_zayin_old, km * 4);
        A((void **)&h->d_kv, (size_t)KV_COUNT * s.K_ld * 4);
        A((void **)&h->d_small, (2 * s.K_ld + 2) * 8);
        A((void **)&h->d_tsum, (2 * (3 * s.K_ld + 1) + 1) * 8);
        for (int eb = 0; eb < 2 && e == cudaSuccess; eb++)
            e = cudaFuncSetAttribute((const void *)kCtpfEstep[s.layout][eb], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem_optin);
        if (e != cudaSuccess) rc = fail((int)e, "device allocation failed: %s", cudaGetErrorString(e));
    }
    if (rc != 0) {
        ctpf_free(h);
        delete h;
        return rc;
    }
    for (auto *v : {&h->bet, &h->vav, &h->dalet, &h->het, &h->bet_old, &h->vav_old, &h->dalet_old, &h->het_old}) v->assign(K, 1.0);  // gpuCTPF.jl:110-119
    h->h_small.assign(2 * K_ld + 2, 0.0);
    h->h_tsum.assign(2 * (3 * K_ld + 1), 0.0);
    *out = h;
    return 0;
}

int tmvb_ctpf_destroy(tmvb_ctpf_t h)
{
    if (!h) return 0;
    ctpf_free(h);
    delete h;
    return 0;
}

int tmvb_ctpf_set_corpus(tmvb_ctpf_t h, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts, const int64_t *R_cumsum,
                         const int64_t *readers, const int64_t *ratings)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    const size_t per_tok = (size_t)s.RS * 4 + 8;
    const int cap2_max = 64;
    // at most 64 term rows staged per document: more resident documents per SM beat shared-memory reads of the remaining
    // rows (CiteULike K=30: 0.70 ms with full tiles, 0.52 ms with 64-token tiles; 16-token tiles 0.68 ms)
    s.tile_cap_max = 64;
    TMVB_TRY(shard_set_corpus(&s, N_cumsum, terms, counts, ctpf_fixed_smem(s.RS, s.lpt) + cap2_max * per_tok));
    TMVB_TRY(shard_pack_aux(&s, R_cumsum, readers, ratings, h->U, &h->d_r_off, &h->d_readers, &h->d_ratings, &h->nnz_r, &h->r_len));
    // size the reader tile of each launch for (about) the 90th percentile of its documents' reader counts;
    // longer reader lists read their overflow rows from L2
    for (Bucket &b : s.buckets) {
        std::vector<int> rl(h->r_len.begin() + b.doc_begin, h->r_len.begin() + b.doc_end);
        std::sort(rl.begin(), rl.end());
        const int p90 = rl.empty() ? 0 : rl[(rl.size() - 1) * 9 / 10];
        b.cap2 = std::min(cap2_max, std::max(16, (p90 + 15) / 16 * 16));
        b.smem = ctpf_fixed_smem(s.RS, s.lpt) + (size_t)(b.cap + b.cap2) * per_tok;
        b.grid = 0;
    }
    h->readers_set = true;
    return 0;
}

int tmvb_ctpf_upload(tmvb_ctpf_t h, const double *hyp, const float *alef, const float *he, const float *bet, const float *vav,
                     const float *gimel, const float *zayin, const float *dalet, const float *het)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    const int K = (int)s.K;
    static const char *hn = "abcdefgh";
    if (hyp)
        for (int q = 0; q < 8; q++) {
            if (!(hyp[q] > 0.0)) return fail(-5, "%c must be positive.", hn[q]);  // modelutils.jl:319-326
            h->hyp[q] = hyp[q];
        }
    auto setk = [&](const float *src, std::vector<double> &dst, std::vector<double> &old, const char *name) -> int {
        if (!src) return 0;
        for (int i = 0; i < K; i++) {
            if (!isfinite(src[i])) return fail(-5, "%s must be finite.", name);
            if (!(src[i] > 0.f)) return fail(-5, "%s must be positive.", name);
            dst[i] = old[i] = (double)src[i];  // the *_old copies start equal (CTPF.jl:89-99)
        }
        return 0;
    };
    TMVB_TRY(setk(bet, h->bet, h->bet_old, "bet"));
    TMVB_TRY(setk(vav, h->vav, h->vav_old, "vav"));
    TMVB_TRY(setk(dalet, h->dalet, h->dalet_old, "dalet"));
    TMVB_TRY(setk(het, h->het, h->het_old, "het"));
    if (alef && s.V > 0) {
        TMVB_TRY(shard_upload_rows(&s, alef, s.d_stats, s.V, nullptr, 2));
        TMVB_TRY(ctpf_build_table(h, 0, 0.0f, 1));
        TMVB_CUDA(cudaMemcpyAsync(s.d_beta[s.cur ^ 1], s.d_beta[s.cur], (size_t)s.V * s.K_ld * 4, cudaMemcpyDeviceToDevice, s.stream));
        TMVB_CUDA(cudaMemcpyAsync(h->d_alef_old, h->d_alef, (size_t)s.V * s.K_ld * 4, cudaMemcpyDeviceToDevice, s.stream));
    }
    if (he && h->U > 0) {
        const int64_t Ksave = s.K;
        (void)Ksave;
        // shard_upload_rows is row-count agnostic: U rows of K floats
        TMVB_TRY(shard_upload_rows(&s, he, h->d_hstats, h->U, nullptr, 2));
        TMVB_TRY(ctpf_build_table(h, 1, 0.0f, 1));
        TMVB_CUDA(cudaMemcpyAsync(h->d_H[h->hcur ^ 1], h->d_H[h->hcur], (size_t)h->U * s.K_ld * 4, cudaMemcpyDeviceToDevice, s.stream));
        TMVB_CUDA(cudaMemcpyAsync(h->d_he_old, h->d_he, (size_t)h->U * s.K_ld * 4, cudaMemcpyDeviceToDevice, s.stream));
    }
    if ((gimel || zayin) && s.M > 0) {
        TMVB_CHECK_ARG(s.corpus_set, "set_corpus must precede the upload of per-document parameters");
        TMVB_TRY(shard_upload_rows(&s, gimel, h->d_gimel, s.M, s.d_perm, 2));
        if (gimel) TMVB_CUDA(cudaMemcpyAsync(h->d_gimel_old, h->d_gimel, (size_t)s.M * s.K_ld * 4, cudaMemcpyDeviceToDevice, s.stream));
        TMVB_TRY(shard_upload_rows(&s, zayin, h->d_zayin, s.M, s.d_perm, 2));
        if (zayin) TMVB_CUDA(cudaMemcpyAsync(h->d_zayin_old, h->d_zayin, (size_t)s.M * s.K_ld * 4, cudaMemcpyDeviceToDevice, s.stream));
    }
    int verr = 0;
    TMVB_TRY(shard_validation(&s, &verr));
    if (verr & 0x10) return fail(-5, "alef, he, gimel and zayin must be finite.");   // modelutils.jl:328-349
    if (verr & 0x20) return fail(-5, "alef, he, gimel and zayin must be positive.");
    TMVB_TRY(ctpf_fetch_tsum(h));
    TMVB_TRY(ctpf_push_kv(h));
    h->elbo_valid = false;
    return 0;
}

int tmvb_ctpf_estep(tmvb_ctpf_t h, int viter, float vtol, int want_elbo)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CHECK_ARG(viter >= 1, "viter must be at least 1");
    TMVB_CHECK_ARG(vtol >= 0.f, "tolerance parameters must be nonnegative");  // gpuCTPF.jl:679
    Shard &s = h->s;
    TMVB_CHECK_ARG(s.corpus_set && h->readers_set, "set_corpus has not been called");
    TMVB_CUDA(cudaSetDevice(s.device));
    CtpfDev p = ctpf_view(h);
    p.viter = viter;
    p.vtol = vtol;
    TMVB_CUDA(cudaEventRecord(s.ev[0], s.stream));
    TMVB_CUDA(cudaMemsetAsync(h->d_small, 0, (2 * s.K_ld + 2) * 8, s.stream));
    const void *fns[2] = {(const void *)kCtpfEstep[s.layout][want_elbo != 0], (const void *)kCtpfEstep[s.layout][want_elbo != 0]};
    TMVB_TRY(shard_launch(&s, pick_by_warps, fns, &p, sizeof(p)));
    TMVB_CUDA(cudaEventRecord(s.ev[1], s.stream));
    s.estep_timed = true;
    h->elbo_valid = (want_elbo != 0);
    return 0;
}

int tmvb_ctpf_reduce_buffers(tmvb_ctpf_t h, void **stats_alef, int64_t *n_alef, void **stats_he, int64_t *n_he, void **small, int64_t *n_small)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    if (stats_alef) *stats_alef = h->s.d_stats;
    if (n_alef) *n_alef = (int64_t)h->s.V * h->s.K_ld;
    if (stats_he) *stats_he = h->d_hstats;
    if (n_he) *n_he = h->U * h->s.K_ld;
    if (small) *small = h->d_small;
    if (n_small) *n_small = 2 * h->s.K_ld + 2;
    return 0;
}

static PeerReduce ctpf_peer_bufs(tmvb_ctpf_t h)
{
    PeerReduce b;
    b.f[0] = h->s.d_stats;
    b.nf[0] = (long long)h->s.V * h->s.K_ld;
    if (h->U > 0) {
        b.f[1] = h->d_hstats;
        b.nf[1] = (long long)h->U * h->s.K_ld;
    }
    b.small = h->d_small;
    b.n_small = 2 * h->s.K_ld + 2;
    return b;
}

/* ---- multi-GPU: the statistics summed over the ranks by ONE kernel over CUDA-IPC peer memory (tmvb_peer.cu) instead of one NCCL
 * all-reduce per buffer.  Handshake as for gpuLDA: export -> all-gather the blobs over any transport -> connect; then
 * tmvb_ctpf_peer_reduce(h) between estep and mstep on every rank. ---- */
int tmvb_ctpf_comm_export(tmvb_ctpf_t h, void *blob, int64_t blob_bytes)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CHECK_ARG(blob_bytes >= TMVB_COMM_BLOB_BYTES, "blob must hold TMVB_COMM_BLOB_BYTES");
    TMVB_CUDA(cudaSetDevice(h->s.device));
    return peer_export(&h->comm, ctpf_peer_bufs(h), blob, (size_t)blob_bytes);
}

int tmvb_ctpf_comm_connect(tmvb_ctpf_t h, int rank, int world, const void *blobs, int64_t blob_bytes)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CUDA(cudaSetDevice(h->s.device));
    TMVB_CUDA(cudaStreamSynchronize(h->s.stream));
    return comm_connect(&h->comm, rank, world, blobs, (size_t)blob_bytes);
}

int tmvb_ctpf_peer_reduce(tmvb_ctpf_t h)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CUDA(cudaSetDevice(h->s.device));
    TMVB_TRY(peer_allreduce(&h->comm, ctpf_peer_bufs(h), h->s.stream, h->s.n_sm));
    h->s.st.kernel_launches++;
    return 0;
}

int tmvb_ctpf_mstep(tmvb_ctpf_t h, int64_t M_total)
{
    TMVB_CHECK_ARG(h != nullptr && M_total > 0, "bad arguments");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    TMVB_CUDA(cudaEventRecord(s.ev[2], s.stream));
    const int K = (int)s.K, K_ld = s.K_ld;
    const double b = h->hyp[1], dd = h->hyp[3], f = h->hyp[5], hh = h->hyp[7];
    // update_he!() / update_alef!() (CTPF.jl:251-270): he_old <- he ; he <- e + statistics  (likewise alef)
    std::swap(h->d_alef, h->d_alef_old);
    std::swap(h->d_he, h->d_he_old);
    s.cur ^= 1;
    h->hcur ^= 1;
    TMVB_TRY(ctpf_build_table(h, 0, (float)h->hyp[0], 1));
    TMVB_TRY(ctpf_build_table(h, 1, (float)h->hyp[4], 1));
    TMVB_CUDA(cudaMemcpyAsync(s.h_pinned + 2 * (3 * K_ld + 1), h->d_small, (2 * K_ld + 2) * 8, cudaMemcpyDeviceToHost, s.stream));
    TMVB_TRY(ctpf_fetch_tsum(h));
    memcpy(h->h_small.data(), s.h_pinned + 2 * (3 * K_ld + 1), (2 * K_ld + 2) * 8);
    if (h->comm.connected) {
        int pst = 0;
        TMVB_TRY(peer_status(&h->comm, s.stream, &pst));
    }
    s.st.d2h_bytes += (2 * K_ld + 2) * 8;
    const double *AL = h->h_tsum.data() + K_ld, *HE = h->h_tsum.data() + (3 * K_ld + 1) + K_ld;
    const double *Gs = h->h_small.data(), *Zs = Gs + K_ld;
    h->dalet_old = h->dalet;
    h->het_old = h->het;
    h->bet_old = h->bet;
    h->vav_old = h->vav;
    for (int i = 0; i < K; i++) {
        h->dalet[i] = dd + AL[i] / h->bet_old[i] + HE[i] / h->vav_old[i];  // update_dalet! CTPF.jl:295-298
        h->het[i] = hh + HE[i] / h->vav_old[i];                            // update_het!   CTPF.jl:302-305
        h->bet[i] = b + Gs[i] / h->dalet[i];                               // update_bet!   CTPF.jl:281-284
        h->vav[i] = f + Gs[i] / h->dalet[i] + Zs[i] / h->het[i];           // update_vav!   CTPF.jl:288-291
    }
    TMVB_TRY(ctpf_push_kv(h));
    TMVB_CUDA(cudaEventRecord(s.ev[3], s.stream));
    s.mstep_timed = true;
    return 0;
}

int tmvb_ctpf_elbo(tmvb_ctpf_t h, int mode, int64_t M_total, double *elbo_docs, double *elbo_global)
{
    TMVB_CHECK_ARG(h && elbo_docs && elbo_global, "NULL argument");
    TMVB_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 or 1");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    const int K = (int)s.K, K_ld = s.K_ld;
    const double a = h->hyp[0], b = h->hyp[1], c = h->hyp[2], dd = h->hyp[3], e = h->hyp[4], f = h->hyp[5], g = h->hyp[6], hh = h->hyp[7];
    const double Md = (double)M_total, Vd = (double)s.V, Ud = (double)h->U;
    const double *T1a = h->h_tsum.data(), *AL = T1a + K_ld, *T3a = AL + K_ld, T4a = h->h_tsum[3 * K_ld];
    const double *T1h = h->h_tsum.data() + (3 * K_ld + 1), *HE = T1h + K_ld, *T3h = HE + K_ld, T4h = h->h_tsum[(3 * K_ld + 1) + 3 * K_ld];
    // Elogpbeta - Elogqbeta + Elogpeta - Elogqeta (CTPF.jl:143-168,197-221) from the table sums
    double glob = Vd * K * (a * log(b) - lgamma(a)) + Ud * K * (e * log(f) - lgamma(e));
    for (int i = 0; i < K; i++) {
        glob += (a - 1.0) * (T1a[i] - Vd * log(h->bet[i])) - b * AL[i] / h->bet[i] + AL[i] - Vd * log(h->bet[i]) + T3a[i];
        glob += (e - 1.0) * (T1h[i] - Ud * log(h->vav[i])) - f * HE[i] / h->vav[i] + HE[i] - Ud * log(h->vav[i]) + T3h[i];
    }
    if (mode == 0) {
        TMVB_CHECK_ARG(h->elbo_valid, "mode 0 needs estep(want_elbo=1) followed by mstep");
        const double *Gs = h->h_small.data(), *Zs = Gs + K_ld;
        double x = T4a + T4h;  // sum S psi(alef), sum S psi(he): the table halves of Elogpz / Elogpya / Elogpyb
        for (int i = 0; i < K; i++) {
            const double lnd = log(h->dalet[i]), lnb = log(h->bet[i]), lnv = log(h->vav[i]), lnh = log(h->het[i]);
            const double Sphi = AL[i] - Vd * a, Sxa = Gs[i] - Md * c - Sphi, Sxb = Zs[i] - Md * g;
            x -= Gs[i] * HE[i] / (h->dalet[i] * h->vav[i]) + Zs[i] * HE[i] / (h->het[i] * h->vav[i]) + Gs[i] * AL[i] / (h->dalet[i] * h->bet[i]);
            x += Sxa * (-lnd - lnv) + Sxb * (-lnh - lnv) + Sphi * (-lnd - lnb);
            x += -(c - 1.0) * Md * lnd - dd * Gs[i] / h->dalet[i] - (g - 1.0) * Md * lnh - hh * Zs[i] / h->het[i];
            x += Gs[i] - Md * lnd + Zs[i] - Md * lnh;
        }
        x += Md * K * (c * log(dd) - lgamma(c)) + Md * K * (g * log(hh) - lgamma(g));
        *elbo_docs = h->h_small[2 * K_ld];
        *elbo_global = glob + x;
        return 0;
    }
    TMVB_CHECK_ARG(s.corpus_set && h->readers_set, "set_corpus has not been called");
    CtpfDev p = ctpf_view(h);
    double *out = h->d_tsum + 2 * (3 * K_ld + 1);
    TMVB_CUDA(cudaMemsetAsync(out, 0, 8, s.stream));
    if (s.M > 0) {
        const int grid = grid_for(s.M * 32, 128, s.n_sm);
#define TMVB_CTPF_ELBO(R) ctpf_elbo_kernel<R><<<grid, 128, 0, s.stream>>>(p, s.d_beta[s.cur ^ 1], h->d_H[h->hcur ^ 1], h->d_alef, h->d_he, (float)dd, (float)hh, out)
        if (s.K <= 32) TMVB_CTPF_ELBO(1); else if (s.K <= 64) TMVB_CTPF_ELBO(2); else if (s.K <= 128) TMVB_CTPF_ELBO(4); else TMVB_CTPF_ELBO(8);
#undef TMVB_CTPF_ELBO
        TMVB_CUDA(cudaGetLastError());
        s.st.kernel_launches++;
    }
    TMVB_CUDA(cudaMemcpyAsync(s.h_pinned, out, 8, cudaMemcpyDeviceToHost, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    s.st.d2h_bytes += 8;
    *elbo_docs = s.h_pinned[0] + (double)s.M * K * (c * log(dd) - lgamma(c) + g * log(hh) - lgamma(g));
    *elbo_global = glob;
    return 0;
}

int tmvb_ctpf_download(tmvb_ctpf_t h, float *alef, float *he, float *bet, float *vav, float *gimel, float *zayin, float *dalet, float *het)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    const int K = (int)s.K;
    for (int i = 0; i < K; i++) {
        if (bet) bet[i] = (float)h->bet[i];
        if (vav) vav[i] = (float)h->vav[i];
        if (dalet) dalet[i] = (float)h->dalet[i];
        if (het) het[i] = (float)h->het[i];
    }
    TMVB_TRY(shard_download_rows(&s, h->d_alef, alef, s.V, nullptr));
    TMVB_TRY(shard_download_rows(&s, h->d_he, he, h->U, nullptr));
    TMVB_TRY(shard_download_rows(&s, h->d_gimel, gimel, s.M, s.d_perm));
    TMVB_TRY(shard_download_rows(&s, h->d_zayin, zayin, s.M, s.d_perm));
    return 0;
}

int tmvb_ctpf_download_old(tmvb_ctpf_t h, float *alef_old, float *he_old, float *bet_old, float *vav_old, float *gimel_old, float *zayin_old,
                           float *dalet_old, float *het_old)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    const int K = (int)s.K;
    for (int i = 0; i < K; i++) {
        if (bet_old) bet_old[i] = (float)h->bet_old[i];
        if (vav_old) vav_old[i] = (float)h->vav_old[i];
        if (dalet_old) dalet_old[i] = (float)h->dalet_old[i];
        if (het_old) het_old[i] = (float)h->het_old[i];
    }
    TMVB_TRY(shard_download_rows(&s, h->d_alef_old, alef_old, s.V, nullptr));
    TMVB_TRY(shard_download_rows(&s, h->d_he_old, he_old, h->U, nullptr));
    TMVB_TRY(shard_download_rows(&s, h->d_gimel_old, gimel_old, s.M, s.d_perm));
    TMVB_TRY(shard_download_rows(&s, h->d_zayin_old, zayin_old, s.M, s.d_perm));
    return 0;
}

/* topics = per-topic ranking of Ebeta = alef ./ bet (gpuCTPF.jl:706-707): scaling a row by 1/bet_i does not change its order */
int tmvb_ctpf_topics(tmvb_ctpf_t h, int32_t *topics)
{
    TMVB_CHECK_ARG(h && topics, "NULL argument");
    TMVB_CUDA(cudaSetDevice(h->s.device));
    return shard_topics(&h->s, h->d_alef, nullptr, topics);
}

// exclusive prefix sums of (len - nmask[seg]) -- one thread: nseg is the number of users or documents
__global__ void recs_offsets_kernel(const int *__restrict__ nmask, int nseg, int len, long long *__restrict__ off)
{
    long long a = 0;
    for (int s = 0; s < nseg; s++) {
        off[s] = a;
        a += len - nmask[s];
    }
    off[nseg] = a;
}
__global__ void recs_nmask_d_kernel(const long long *__restrict__ r_off, const int *__restrict__ perm, long long M, int *__restrict__ nmask_d)
{
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < M; p += (long long)gridDim.x * blockDim.x)
        nmask_d[perm[p]] = (int)(r_off[p + 1] - r_off[p]);
}

/* scores / urecs / drecs of train!(::gpuCTPF) (gpuCTPF.jl:709-731) on the device -- see tmvb_recs.cuh.  Every output is optional:
 *   scores  [M x U] column-major Float32 (Julia's model.scores)
 *   urecs   concatenated rankings, urecs[uoff[u] .. uoff[u+1]) = documents (1-based) not in user u's library, by descending score
 *   drecs   concatenated rankings, drecs[doff[d] .. doff[d+1]) = users (1-based) who have not read document d, by descending score
 *   uoff [U + 1], doff [M + 1]  (both rankings have M * U - sum(R) entries)
 * mode bit 0: contraction on the CUDA cores in fp32 instead of the tensor cores (the checker of the tcgen05 kernel). */
int tmvb_ctpf_recs(tmvb_ctpf_t h, float *scores, int32_t *urecs, int64_t *uoff, int32_t *drecs, int64_t *doff, int mode)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CHECK_ARG(s.corpus_set && h->readers_set, "set_corpus has not been called");
    TMVB_CHECK_ARG((urecs == nullptr) == (uoff == nullptr) && (drecs == nullptr) == (doff == nullptr), "a ranking needs its offsets array");
    TMVB_CUDA(cudaSetDevice(s.device));
    const int64_t M = s.M, U = h->U;
    const int K = (int)s.K, KP = (K + kRecBK - 1) / kRecBK * kRecBK;
    if (M == 0 || U == 0) {
        if (uoff) memset(uoff, 0, (U + 1) * 8);
        if (doff) memset(doff, 0, (M + 1) * 8);
        return 0;
    }
    const int ld_d = (int)((U + 3) / 4 * 4), ld_u = (int)((M + 3) / 4 * 4);
    const bool want_u = urecs != nullptr || scores != nullptr, want_d = drecs != nullptr;
    float *X = nullptr, *Y = nullptr, *kv = nullptr, *keys_d = nullptr, *keys_u = nullptr;
    int *nmask = nullptr, *out = nullptr;
    long long *off = nullptr;
    void *ws = nullptr;
    size_t ws_bytes = 0;
    int rc = 0;
    auto cleanup = [&]() {
        cudaStreamSynchronize(s.stream);
        cudaFree(X);
        cudaFree(Y);
        cudaFree(kv);
        cudaFree(keys_d);
        cudaFree(keys_u);
        cudaFree(nmask);
        cudaFree(out);
        cudaFree(off);
        cudaFree(ws);
    };
#define RECS_CUDA(expr)                                                                                              \
    do {                                                                                                             \
        cudaError_t _e = (expr);                                                                                     \
        if (_e != cudaSuccess) {                                                                                     \
            cleanup();                                                                                               \
            return fail((int)_e, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e));      \
        }                                                                                                            \
    } while (0)
    RECS_CUDA(cudaMalloc((void **)&X, (size_t)M * KP * 4));
    RECS_CUDA(cudaMalloc((void **)&Y, (size_t)U * KP * 4));
    RECS_CUDA(cudaMalloc((void **)&kv, (size_t)3 * K * 4));
    RECS_CUDA(cudaMalloc((void **)&nmask, (size_t)(M + U) * 4));
    RECS_CUDA(cudaMalloc((void **)&off, (size_t)(M + U + 2) * 8));
    RECS_CUDA(cudaMalloc((void **)&out, (size_t)std::max<int64_t>(M * U - h->nnz_r, 1) * 4));
    if (want_d) RECS_CUDA(cudaMalloc((void **)&keys_d, (size_t)M * ld_d * 4));
    if (want_u) RECS_CUDA(cudaMalloc((void **)&keys_u, (size_t)U * ld_u * 4));
    {   // Eeta = he ./ vav, Etheta = gimel ./ dalet, Eepsilon = zayin ./ het (gpuCTPF.jl:709-713) from the fp64 masters of the rates
        std::vector<float> hv(3 * K);
        for (int i = 0; i < K; i++) {
            hv[i] = (float)(1.0 / h->dalet[i]);
            hv[K + i] = (float)(1.0 / h->het[i]);
            hv[2 * K + i] = (float)(1.0 / h->vav[i]);
        }
        RECS_CUDA(cudaMemcpyAsync(kv, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice, s.stream));
        RECS_CUDA(cudaStreamSynchronize(s.stream));
        s.st.h2d_bytes += hv.size() * 4;
    }
    recs_theta_kernel<<<grid_for(M * KP, 256, s.n_sm), 256, 0, s.stream>>>(h->d_gimel, h->d_zayin, kv, kv + K, s.d_perm, M, K, s.K_ld, KP, X);
    recs_eta_kernel<<<grid_for(U * KP, 256, s.n_sm), 256, 0, s.stream>>>(h->d_he, kv + 2 * K, U, K, s.K_ld, KP, Y);
    RECS_CUDA(cudaGetLastError());
    const size_t smem = (size_t)(2 * kRecBM + 2 * kRecBN) * 128;
    RECS_CUDA(cudaFuncSetAttribute((const void *)recs_scores_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto contract = [&](const float *A, const float *B, float *Cm, int P, int Q, int ldc, int swap) {
        if (mode & 1) {
            recs_scores_ref_kernel<<<grid_for((long long)P * Q, 256, s.n_sm), 256, 0, s.stream>>>(A, B, Cm, P, Q, KP, ldc);
        } else {
            const int gx = (P + kRecBM - 1) / kRecBM, ntile = (Q + kRecBN - 1) / kRecBN;
            const int gy = std::max(1, std::min(ntile, (2 * s.n_sm + gx - 1) / gx));
            recs_scores_umma_kernel<<<dim3(gx, gy), 128, smem, s.stream>>>(A, B, Cm, P, Q, KP, ldc, swap);
        }
        s.st.kernel_launches++;
    };
    if (want_d) contract(X, Y, keys_d, (int)M, (int)U, ld_d, 0);
    if (want_u) contract(Y, X, keys_u, (int)U, (int)M, ld_u, 1);
    RECS_CUDA(cudaGetLastError());
    if (scores) {
        RECS_CUDA(cudaMemcpy2DAsync(scores, (size_t)M * 4, keys_u, (size_t)ld_u * 4, (size_t)M * 4, (size_t)U, cudaMemcpyDeviceToHost, s.stream));
        s.st.d2h_bytes += M * U * 4;
    }
    if (urecs || drecs) {
        int *nmask_d = nmask, *nmask_u = nmask + M;
        RECS_CUDA(cudaMemsetAsync(nmask, 0, (size_t)(M + U) * 4, s.stream));
        recs_nmask_d_kernel<<<grid_for(M, 256, s.n_sm), 256, 0, s.stream>>>(h->d_r_off, s.d_perm, M, nmask_d);
        recs_mask_kernel<<<grid_for(M * 32, 256, s.n_sm), 256, 0, s.stream>>>(h->d_r_off, h->d_readers, s.d_perm, M, drecs ? keys_d : nullptr, ld_d,
                                                                          keys_u, ld_u, nmask_u);
        RECS_CUDA(cudaGetLastError());
        long long *off_d = off, *off_u = off + M + 1;
        const int64_t total = M * U - h->nnz_r;
        if (drecs) {
            recs_offsets_kernel<<<1, 1, 0, s.stream>>>(nmask_d, (int)M, (int)U, off_d);
            rc = segmented_rank(keys_d, (int)M, (int)U, ld_d, nmask_d, off_d, out, &ws, &ws_bytes, s.stream, s.n_sm);
            if (rc) {
                cleanup();
                return rc;
            }
            RECS_CUDA(cudaMemcpyAsync(drecs, out, (size_t)total * 4, cudaMemcpyDeviceToHost, s.stream));
            RECS_CUDA(cudaMemcpyAsync(doff, off_d, (size_t)(M + 1) * 8, cudaMemcpyDeviceToHost, s.stream));
            RECS_CUDA(cudaStreamSynchronize(s.stream));
            s.st.d2h_bytes += total * 4 + (M + 1) * 8;
        }
        if (urecs) {
            recs_offsets_kernel<<<1, 1, 0, s.stream>>>(nmask_u, (int)U, (int)M, off_u);
            rc = segmented_rank(keys_u, (int)U, (int)M, ld_u, nmask_u, off_u, out, &ws, &ws_bytes, s.stream, s.n_sm);
            if (rc) {
                cleanup();
                return rc;
            }
            RECS_CUDA(cudaMemcpyAsync(urecs, out, (size_t)total * 4, cudaMemcpyDeviceToHost, s.stream));
            RECS_CUDA(cudaMemcpyAsync(uoff, off_u, (size_t)(U + 1) * 8, cudaMemcpyDeviceToHost, s.stream));
            s.st.d2h_bytes += total * 4 + (U + 1) * 8;
        }
        s.st.kernel_launches += 8;
    }
    cleanup();
#undef RECS_CUDA
    return 0;
}

int tmvb_ctpf_get_stats(tmvb_ctpf_t h, tmvb_stats *out)
{
    TMVB_CHECK_ARG(h && out, "NULL argument");
    return shard_get_stats(&h->s, h->d_small + 2 * h->s.K_ld + 1, out);
}

}  // extern "C"
