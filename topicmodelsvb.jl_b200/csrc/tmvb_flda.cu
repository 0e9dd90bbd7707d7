// tmvb_flda.cu -- filtered LDA (src/fLDA.jl) on the device: the model the reference's `@gpu` macro skips (macros.jl:274-275,
// "elseif isa(model, fLDA) nothing") and its todo list asks for.  Same shard plumbing as the other
// handles (tmvb_shard.cuh); the per-document inner loop of train!(::fLDA) (fLDA.jl:223-233) is ONE kernel per length bucket.
//
// Per sweep and token n of document d (fLDA.jl:175-201), with tau_n in [0, 1] the token's "topical" probability:
//     phi_ni  = softmax_i( tau_n ln(beta[i, w_n] + eps) + Elogtheta_di )                              update_phi!
//     tau_n   = eta / (eps + eta + (1 - eta) kappa[w_n] prod_i beta[i, w_n]^(-phi_ni))                update_tau!
//     gamma_d = eps + alpha + phi c ,   Elogtheta_d = psi(gamma_d) - psi(sum gamma_d)                 update_gamma!, update_Elogtheta!
// Device form: the K x V table holds L = log2(beta + eps) (rebuilt once per M-step), the document vector is
// E2_i = (Elogtheta_di - max_i Elogtheta_di) log2 e, so that p_ni = 2^(tau_n L_ni + E2_i) needs no per-token maximum
// (L <= 0, and the topic with E2 = 0 keeps s_n >= 2^-99 > FLT_MIN); one FFMA2 + two MUFU.EX2 per topic pair, then
// s_n = sum_i p_ni, q_n = sum_i p_ni L_ni (the exponent of the product in update_tau!) and g += p_n c_n / s_n.
// Where the reference's update_tau! raises beta WITHOUT epsilon (fLDA.jl:193) the table's log2(beta + eps) is used: the two
// differ only for beta < 2^-76, where phi_ni is itself below 2^-70 of the column.
// tau / tau_old persist per token across outer iterations (fLDA.jl:50-51): [nnz] arrays in the shard's internal order, staged
// into shared memory with the document's rows.  The scatter pass recomputes the last phi from (tau_old, E2 of the last sweep)
// and adds phi tau c into the K x V statistics (update_beta!(model, d), fLDA.jl:168-171) and (1 - tau) c into the V-vector
// of update_kappa!(model, d) (fLDA.jl:156-159).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "tmvb_comm.cuh"
#include "tmvb_filt.cuh"
#include "tmvb_shard.cuh"

namespace tmvb {

struct FldaDev {
    int K, K_ld, V, RS;
    long long M;
    const float *L;       // [V][K_ld] log2(beta + eps); pad topics 0
    const float *kq;      // [V] (1 - eta) kappa_j, then [V] = eta itself (device-resident so that the launch graph does not change with eta)
    const float *alpha;   // [K_ld]
    float *stats;         // [V][K_ld]
    float *kstats;        // [V]
    const long long *doc_off;
    const int *terms;
    const float *counts;
    const float *doc_c;
    float *Elogtheta, *Elogtheta_old, *gamma;   // [M][K_ld]
    float *tau, *tau_old;                        // [nnz]
    float *doc_tc;                               // [M] sum_n tau_n c_n
    double *small;                               // [K_ld] sum_d Elogtheta_d | [K_ld] sweeps | [K_ld + 1] sum_dn tau_n c_n
    int viter;
    float vtol;
    int dbg;   // bit 0: no scatter (predict), bit 1: skip the final pass (developer probe)
};

static size_t flda_fixed_smem(int RS, int lpt) { return 64 + (size_t)(32 / lpt) * RS * 4 + (size_t)RS * 4; }
constexpr size_t kFldaPerTokExtra = 12;   // tau_s, tauo_s, kq_s

// One warp per document.  Token phase: lane (ts = lane / LPT, kl = lane % LPT) owns token stream ts and the 16-byte chunks
// kl + LPT m of every row; K phase: lane l owns topics l + 32 r (tmvb_estep.cuh).
template <int LPT, int CPL>
__global__ void __launch_bounds__(32) flda_estep_kernel(const FldaDev p, int doc_begin, int doc_end, int cap, int cap2, int *counter)
{
    constexpr int S = 32 / LPT;
    constexpr int R = (4 * LPT * CPL + 31) / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x, kl = lane % LPT, ts = lane / LPT;
    const int K = p.K, K_ld = p.K_ld, CH = K_ld >> 2, RS = p.RS;
    (void)cap2;
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw);
    float *gs = reinterpret_cast<float *>(smem_raw + 64);   // [S][RS]
    float *e_s = gs + (size_t)S * RS;                        // [RS]
    float *tile = e_s + RS;                                  // [cap][RS]
    float *cnt_s = tile + (size_t)cap * RS;                  // [cap]
    int *term_s = reinterpret_cast<int *>(cnt_s + cap);      // [cap]
    float *tau_s = reinterpret_cast<float *>(term_s + cap);  // [cap]
    float *tauo_s = tau_s + cap;                             // [cap]
    float *kq_s = tauo_s + cap;                              // [cap]

    float alpha_k[R], Eold_k[R], Enew_k[R], gam_k[R];
    double esum_k[R];
    float asum = 0.0f;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = lane + 32 * r;
        alpha_k[r] = (i < K) ? p.alpha[i] : 0.0f;
        asum += alpha_k[r];
        esum_k[r] = 0.0;
        Eold_k[r] = Enew_k[r] = gam_k[r] = 0.0f;
    }
    asum = warp_sum(asum);
    const float eta = __ldg(p.kq + p.V);
    double tc_thr = 0.0;
    unsigned long long sweeps_thr = 0;
    unsigned phase = 0;
    if (lane == 0) mbar_init(mbar, 1);
    __syncwarp();

    const int chunk = max(1, min(8, (doc_end - doc_begin) / (4 * (int)gridDim.x)));
    int d_next = 0, d_lim = 0;
    for (;;) {
        if (d_next >= d_lim) {
            if (lane == 0) d_next = doc_begin + atomicAdd(counter, chunk);
            d_next = __shfl_sync(0xffffffffu, d_next, 0);
            d_lim = min(d_next + chunk, doc_end);
        }
        if (d_next >= doc_end) break;
        const int d = d_next++;
        const long long o = p.doc_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o);
        const int ns = min(Nd, cap);
        const int rounds = (Nd + S - 1) / S;

        for (int n = lane; n < ns; n += 32) {
            const int term = p.terms[o + n];
            term_s[n] = term;
            cnt_s[n] = p.counts[o + n];
            tau_s[n] = p.tau[o + n];
            tauo_s[n] = p.tau_old[o + n];
            kq_s[n] = __ldg(p.kq + term);
        }
        stage_rows(tile, term_s, p.L, ns, K_ld, RS, lane, mbar, 1);
        float mx = -3.0e38f;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            Eold_k[r] = (i < K) ? p.Elogtheta[(size_t)d * K_ld + i] : 0.0f;
            Enew_k[r] = Eold_k[r];
            if (i < K) mx = fmaxf(mx, Eold_k[r]);
        }
        mx = warp_max(mx);
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            if (i < K_ld) e_s[i] = (i < K) ? (Eold_k[r] - mx) * kLog2e : kPadE;
        }
        // sum(gamma_d) = sum(alpha) + C_d + K eps whatever phi is: psi(sum gamma) (fLDA.jl:177) is a per-document constant
        const float gsum = (asum + __ldg(p.doc_c + d)) + (float)K * TMVB_EPS;
        const float psi_sum = psi_lgamma<false>(gsum).psi;
        stage_wait(mbar, phase, 1);

        int v = 0;
        while (v < p.viter) {
            f32x2 E01[CPL], E23[CPL], g01[CPL], g23[CPL];
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                const bool in = (kl + LPT * m < CH);
                const float4 E = in ? reinterpret_cast<const float4 *>(e_s)[kl + LPT * m] : make_float4(kPadE, kPadE, kPadE, kPadE);
                E01[m] = pk2(E.x, E.y);
                E23[m] = pk2(E.z, E.w);
                g01[m] = g23[m] = 0ull;
            }
            // ---- token phase: update_phi!, update_tau!, and the phi * counts product of update_gamma!
#pragma unroll 2
            for (int r = 0; r < rounds; r++) {
                const int n = r * S + ts;
                const bool ok = n < Nd;
                const int nn = ok ? n : 0;
                const bool in_tile = nn < cap;
                const int term = in_tile ? 0 : __ldg(p.terms + o + nn);
                ulonglong2 b[CPL];
                flda_load_row<LPT, CPL>(tile, p.L, RS, K_ld, CH, nn, cap, term, kl, b);
                const float c = ok ? (in_tile ? cnt_s[nn] : __ldg(p.counts + o + nn)) : 0.0f;
                const float tau = in_tile ? tau_s[nn] : __ldcg(p.tau + o + nn);
                f32x2 p01[CPL], p23[CPL];
                float s, q;
                flda_row<CPL>(b, E01, E23, tau, p01, p23, s, q);
                s = group_sum<LPT>(s);
                q = group_sum<LPT>(q);
                const float rs = rcp_ftz(s);
                const float t = c * rs;
                const f32x2 t2 = pk2(t, t);
#pragma unroll
                for (int m = 0; m < CPL; m++) {
                    g01[m] = fma2(p01[m], t2, g01[m]);
                    g23[m] = fma2(p23[m], t2, g23[m]);
                }
                __syncwarp();   // every lane of the token has read tau before lane kl = 0 replaces it
                if (ok && kl == 0) {
                    // prod_i beta_i^(-phi_i) = 2^(-q / s); eps + eta + (1 - eta) kappa prod  (fLDA.jl:193, @boink on the whole denominator)
                    const float kq = in_tile ? kq_s[nn] : __ldg(p.kq + term);
                    const float den = (eta + kq * ex2_ftz(-q * rs)) + TMVB_EPS;
                    const float tn = fast_div_pos(eta, den);
                    if (in_tile) {
                        tauo_s[nn] = tau;
                        tau_s[nn] = tn;
                    } else {
                        __stcg(p.tau_old + o + nn, tau);
                        __stcg(p.tau + o + nn, tn);
                    }
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
            // ---- K phase: update_gamma! (fLDA.jl:182-185), update_Elogtheta! (fLDA.jl:175-178)
            float dpart = 0.0f, mxn = -3.0e38f;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = lane + 32 * r;
                if (i < K) {
                    const float gi = owner_sum<S>(gs, RS, i);
                    gam_k[r] = (alpha_k[r] + gi) + TMVB_EPS;
                    Eold_k[r] = Enew_k[r];
                    Enew_k[r] = psi_lgamma<false, true>(gam_k[r]).psi - psi_sum;
                    const float df = Enew_k[r] - Eold_k[r];
                    dpart = fmaf(df, df, dpart);
                    mxn = fmaxf(mxn, Enew_k[r]);
                }
            }
            v++;
            __syncwarp();   // the owner sums have been read before gs / e_s are overwritten
            // fLDA.jl:229: stop when ||Elogtheta - Elogtheta_old||_2 < vtol (or after viter sweeps); e_s then still holds the
            // vector the last phi was computed from
            if (v >= p.viter || sqrtf(warp_sum(dpart)) < p.vtol) break;
            mxn = warp_max(mxn);
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = lane + 32 * r;
                if (i < K) e_s[i] = (Enew_k[r] - mxn) * kLog2e;
            }
            __syncwarp();
        }

        // ---- update_beta!(model, d) (fLDA.jl:168-171), update_kappa!(model, d) (fLDA.jl:156-159): the last phi, rebuilt from the
        // tau it was computed from (tau_old) and the e_s of the last sweep; weights tau_n c_n and (1 - tau_n) c_n with the FINAL tau
        float tc = 0.0f;
        if (!(p.dbg & 2) && v > 0) {
            f32x2 E01[CPL], E23[CPL];
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                const bool in = (kl + LPT * m < CH);
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
                const float tauo = in_tile ? tauo_s[nn] : __ldcg(p.tau_old + o + nn);
                const float tauf = in_tile ? tau_s[nn] : __ldcg(p.tau + o + nn);
                f32x2 p01[CPL], p23[CPL];
                float s, q;
                flda_row<CPL>(b, E01, E23, tauo, p01, p23, s, q);
                s = group_sum<LPT>(s);
                if (ok) {
                    const float w = tauf * c * rcp_ftz(s);
                    const f32x2 w2 = pk2(w, w);
                    float *srow = p.stats + (size_t)term * K_ld + 4 * kl;
                    if (!(p.dbg & 1)) {
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
                    if (kl == 0) {
                        tc = fmaf(tauf, c, tc);
                        if (in_tile) {
                            p.tau[o + nn] = tauf;
                            p.tau_old[o + nn] = tauo;
                        }
                    }
                }
            }
        }
        tc = warp_sum(tc);
        if (lane == 0) {
            p.doc_tc[d] = tc;
            tc_thr += (double)tc;
            sweeps_thr += (unsigned long long)v;
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = lane + 32 * r;
            if (i < K_ld) {
                const bool ok = i < K;
                p.gamma[(size_t)d * K_ld + i] = ok ? gam_k[r] : 0.0f;
                p.Elogtheta[(size_t)d * K_ld + i] = ok ? Enew_k[r] : 0.0f;
                p.Elogtheta_old[(size_t)d * K_ld + i] = ok ? Eold_k[r] : 0.0f;
                if (ok) esum_k[r] += (double)Enew_k[r];
            }
        }
        __syncwarp();   // the tile, e_s and the token arrays are free for the next document
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = lane + 32 * r;
        if (i < K && esum_k[r] != 0.0) atomicAdd(p.small + i, esum_k[r]);
    }
    if (lane == 0) {
        if (sweeps_thr) atomicAdd(p.small + K_ld, (double)sweeps_thr);
        if (tc_thr != 0.0) atomicAdd(p.small + K_ld + 1, tc_thr);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// The same inner loop with the per-token scalars in registers ("register-state" variant, LPT = 2: S = 16 token streams).
// flda_estep_kernel keeps tau / tau_old / counts / (1 - eta) kappa of the staged tokens in shared memory and lets lane kl = 0 of a
// token rewrite tau at the end of every round (a warp barrier and a divergent branch per round), and it sizes the tile for the
// longest document of the launch (43-100 KB: 2-5 resident warps per SM at NSF).  Here the tile holds at most TR * S = 64 tokens per
// warp (W = 1 or 2 warps per document, see the kernel; longer documents read the remaining rows from L2 as before), and for its TR
// tile rounds each lane keeps c, tau, tau_old and kq of its token in registers -- both lanes of a token compute the new tau
// redundantly, so the round has no shared-memory write, no barrier and no branch, and two rounds are issued as ONE basic block
// (flda_token2): two independent dependency chains (LDS -> FFMA2 -> MUFU.EX2 -> sums -> SHFL -> MUFU.RCP -> FFMA2) per warp at 12
// resident warps per SM.  Measured on B200 (NSF K=50): E-step 5.23 ms (flda_estep_kernel, full tiles) -> 3.91 (64-token tiles) ->
// 3.55 ms (this kernel); DESIGN.md 4.9.
template <int LPT, int CPL>
__device__ __forceinline__ void flda_tile_row(const float *tile, int RS, int CH, int n, int kl, ulonglong2 (&b)[CPL])
{
    const ulonglong2 zero = make_ulonglong2(0ull, 0ull);
    const ulonglong2 *row = reinterpret_cast<const ulonglong2 *>(tile + (size_t)n * RS) + kl;
#pragma unroll
    for (int m = 0; m < CPL; m++) b[m] = (kl + LPT * m < CH) ? row[LPT * m] : zero;
}

// tau_n = eta / (eta + (1 - eta) kappa prod_i beta_i^(-phi_i) + eps),  prod = 2^(-q / s)   (fLDA.jl:190-195)
__device__ __forceinline__ float flda_new_tau(float eta, float kq, float q, float rs) { return fast_div_pos(eta, (eta + kq * ex2_ftz(-q * rs)) + TMVB_EPS); }

template <int LPT, int CPL>
__device__ __forceinline__ float flda_token1(const float *tile, int RS, int CH, int n, int kl, const f32x2 (&E01)[CPL], const f32x2 (&E23)[CPL], float c,
                                             float tau, float kq, float eta, f32x2 (&g01)[CPL], f32x2 (&g23)[CPL])
{
    ulonglong2 b[CPL];
    flda_tile_row<LPT, CPL>(tile, RS, CH, n, kl, b);
    f32x2 p01[CPL], p23[CPL];
    float s, q;
    flda_row<CPL>(b, E01, E23, tau, p01, p23, s, q);
    s = group_sum<LPT>(s);
    q = group_sum<LPT>(q);
    const float rs = rcp_ftz(s), t = c * rs;
    const f32x2 t2 = pk2(t, t);
#pragma unroll
    for (int m = 0; m < CPL; m++) {
        g01[m] = fma2(p01[m], t2, g01[m]);
        g23[m] = fma2(p23[m], t2, g23[m]);
    }
    return flda_new_tau(eta, kq, q, rs);
}

template <int LPT, int CPL>
__device__ __forceinline__ void flda_token2(const float *tile, int RS, int CH, int na, int nb, int kl, const f32x2 (&E01)[CPL], const f32x2 (&E23)[CPL],
                                            float ca, float cb, float &taua, float &taub, float kqa, float kqb, float eta, f32x2 (&g01)[CPL],
                                            f32x2 (&g23)[CPL])
{
    ulonglong2 ba[CPL], bb[CPL];
    flda_tile_row<LPT, CPL>(tile, RS, CH, na, kl, ba);
    flda_tile_row<LPT, CPL>(tile, RS, CH, nb, kl, bb);
    f32x2 pa01[CPL], pa23[CPL], pb01[CPL], pb23[CPL];
    float sa, qa, sb, qb;
    flda_row<CPL>(ba, E01, E23, taua, pa01, pa23, sa, qa);
    flda_row<CPL>(bb, E01, E23, taub, pb01, pb23, sb, qb);
    sa = group_sum<LPT>(sa);
    sb = group_sum<LPT>(sb);
    qa = group_sum<LPT>(qa);
    qb = group_sum<LPT>(qb);
    const float rsa = rcp_ftz(sa), rsb = rcp_ftz(sb);
    const float ta = ca * rsa, tb = cb * rsb;
    const f32x2 ta2 = pk2(ta, ta), tb2 = pk2(tb, tb);
#pragma unroll
    for (int m = 0; m < CPL; m++) {   // the order of the two additions is the order of the rounds
        g01[m] = fma2(pb01[m], tb2, fma2(pa01[m], ta2, g01[m]));
        g23[m] = fma2(pb23[m], tb2, fma2(pa23[m], ta2, g23[m]));
    }
    taua = flda_new_tau(eta, kqa, qa, rsa);
    taub = flda_new_tau(eta, kqb, qb, rsb);
}

template <int W>
__device__ __forceinline__ void flda_cta_sync()
{
    if (W == 1)
        __syncwarp();
    else
        __syncthreads();
}

// 64-byte header (mbarrier | next-document slot | per-warp partials) | gs [W][S][RS] | En_s [2][RS]
static size_t flda_reg_fixed_smem(int RS, int lpt, int W) { return 64 + (size_t)W * (32 / lpt) * RS * 4 + 2 * (size_t)RS * 4; }
constexpr float kPadRaw = -3.0e4f;   // Elogtheta of a pad topic: (kPadRaw - max) log2(e) flushes 2^x to zero

// W warps share one document and its tile (W = 2 for the launches whose tile exceeds 64 tokens: twice the tile at the same
// shared memory per resident warp, so that NSF's documents fit -- 3 % of the tokens lie beyond 128, 29 % beyond 64): warp w takes the
// rounds w, w + W, ... and the 32-topic slices w, w + W, ... of the K phase.  Two CTA barriers per sweep (after the per-stream
// partial sums are in shared memory; after the new Elogtheta, the partial norms and maxima are), three per document.
template <int LPT, int CPL, int TR, int W, int MAXREG>
__global__ void __maxnreg__(MAXREG) flda_estep_reg_kernel(const FldaDev p, int doc_begin, int doc_end, int cap, int cap2, int *counter)
{
    constexpr int S = 32 / LPT;
    constexpr int R = (4 * LPT * CPL + 31) / 32;   // 32-topic slices of a K vector
    constexpr int RW = (R + W - 1) / W;            // slices per warp: slice warp + W j
    static_assert(W <= 4 && TR % 2 == 0, "header slots / pairing");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, kl = lane % LPT, ts = lane / LPT;
    const int K = p.K, K_ld = p.K_ld, CH = K_ld >> 2, RS = p.RS;
    (void)cap2;
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw);
    int *next_s = reinterpret_cast<int *>(smem_raw + 8);
    float *dpart_s = reinterpret_cast<float *>(smem_raw + 16);   // [W]
    float *mx_s = reinterpret_cast<float *>(smem_raw + 32);      // [W]
    float *tc_s = reinterpret_cast<float *>(smem_raw + 48);      // [W]
    float *gs = reinterpret_cast<float *>(smem_raw + 64);        // [W][S][RS]
    float *En_s = gs + (size_t)W * S * RS;                       // [2][RS]: the Elogtheta sweep v reads is buffer v & 1
    float *tile = En_s + 2 * RS;                                 // [cap][RS],  cap <= W * TR * S, a multiple of S
    int *term_s = reinterpret_cast<int *>(tile + (size_t)cap * RS + cap);   // [cap]  (plan_buckets: rows | 4 bytes | 4 bytes per token | ...)
    float *gs_w = gs + (size_t)warp * S * RS;

    float alpha_k[RW], Eold_k[RW], Enew_k[RW], gam_k[RW];
    double esum_k[RW];
    float asum = 0.0f;
    for (int i = lane; i < K; i += 32) asum += p.alpha[i];
    asum = warp_sum(asum);
#pragma unroll
    for (int j = 0; j < RW; j++) {
        const int i = lane + 32 * (warp + W * j);
        alpha_k[j] = (i < K) ? p.alpha[i] : 0.0f;
        esum_k[j] = 0.0;
        Eold_k[j] = Enew_k[j] = gam_k[j] = 0.0f;
        if (i < K_ld) En_s[i] = En_s[RS + i] = kPadRaw;
    }
    const float eta = __ldg(p.kq + p.V);
    double tc_thr = 0.0;
    unsigned long long sweeps_thr = 0;
    unsigned phase = 0;
    if (tid == 0) mbar_init(mbar, 1);
    flda_cta_sync<W>();

    const int chunk = max(1, min(8, (doc_end - doc_begin) / (4 * (int)gridDim.x)));
    int d_next = 0, d_lim = 0;
    for (;;) {
        if (d_next >= d_lim) {
            if (W == 1) {
                if (lane == 0) d_next = doc_begin + atomicAdd(counter, chunk);
                d_next = __shfl_sync(0xffffffffu, d_next, 0);
            } else {
                if (tid == 0) *next_s = doc_begin + atomicAdd(counter, chunk);
                __syncthreads();
                d_next = *next_s;
            }
            d_lim = min(d_next + chunk, doc_end);
        }
        if (d_next >= doc_end) break;
        const int d = d_next++;
        const long long o = p.doc_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o);
        const int ns = min(Nd, cap);
        const int trounds = (ns + S - 1) / S;     // rounds served by the tile and the register state (<= W * TR)
        const int rounds = (Nd + S - 1) / S;

        // this lane's tokens of its warp's tile rounds: slot j is token (warp + W j) S + ts (an empty slot has c = 0 and reads row 0)
        float c_r[TR], tau_r[TR], tauo_r[TR], kq_r[TR];
        int term_r[TR];
#pragma unroll
        for (int j = 0; j < TR; j++) {
            const int n = (warp + W * j) * S + ts;
            const bool ok = n < ns;
            term_r[j] = ok ? __ldg(p.terms + o + n) : 0;
            c_r[j] = ok ? __ldg(p.counts + o + n) : 0.0f;
            tau_r[j] = ok ? __ldcg(p.tau + o + n) : 0.5f;
            tauo_r[j] = ok ? __ldcg(p.tau_old + o + n) : 0.5f;
        }
#pragma unroll
        for (int j = 0; j < TR; j++) {
            const int n = (warp + W * j) * S + ts;
            kq_r[j] = (n < ns) ? __ldg(p.kq + term_r[j]) : 0.0f;
            if (n < ns && kl == 0) term_s[n] = term_r[j];
        }
        float mx = -3.0e38f;
#pragma unroll
        for (int j = 0; j < RW; j++) {
            const int i = lane + 32 * (warp + W * j);
            Eold_k[j] = (i < K) ? p.Elogtheta[(size_t)d * K_ld + i] : 0.0f;
            Enew_k[j] = Eold_k[j];
            if (i < K) {
                mx = fmaxf(mx, Eold_k[j]);
                En_s[i] = Eold_k[j];
            }
        }
        mx = warp_max(mx);
        if (W > 1 && lane == 0) mx_s[warp] = mx;
        fence_proxy_async_smem();   // the tile was read through the generic proxy; the bulk copies write it through the async proxy
        if (tid == 0) mbar_arrive_expect_tx(mbar, (unsigned)(ns * K_ld * 4));
        flda_cta_sync<W>();
        for (int n = tid; n < ns; n += 32 * W) bulk_g2s(tile + n * RS, p.L + (size_t)term_s[n] * K_ld, (unsigned)(K_ld * 4), mbar);
        if (W > 1) {
            mx = mx_s[0];
#pragma unroll
            for (int w = 1; w < W; w++) mx = fmaxf(mx, mx_s[w]);
        }
        // sum(gamma_d) = sum(alpha) + C_d + K eps whatever phi is: psi(sum gamma) (fLDA.jl:177) is a per-document constant
        const float gsum = (asum + __ldg(p.doc_c + d)) + (float)K * TMVB_EPS;
        const float psi_sum = psi_lgamma<false>(gsum).psi;
        mbar_wait(mbar, phase);
        phase ^= 1u;

        int v = 0;
        for (;;) {
            const float *En = En_s + (v & 1) * RS;
            f32x2 E01[CPL], E23[CPL], g01[CPL], g23[CPL];
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                const bool in = (kl + LPT * m < CH);
                const float4 E = in ? reinterpret_cast<const float4 *>(En)[kl + LPT * m] : make_float4(kPadRaw, kPadRaw, kPadRaw, kPadRaw);
                E01[m] = pk2((E.x - mx) * kLog2e, (E.y - mx) * kLog2e);
                E23[m] = pk2((E.z - mx) * kLog2e, (E.w - mx) * kLog2e);
                g01[m] = g23[m] = 0ull;
            }
            // ---- token phase: update_phi!, update_tau!, and the phi * counts product of update_gamma!
#pragma unroll
            for (int j = 0; j < TR; j += 2) {
                const int ra = warp + W * j, rb = ra + W;
                const int na = ra * S + ts, nb = rb * S + ts;
                if (rb < trounds) {
                    float ta = tau_r[j], tb = tau_r[j + 1];
                    tauo_r[j] = ta;
                    tauo_r[j + 1] = tb;
                    flda_token2<LPT, CPL>(tile, RS, CH, na < ns ? na : 0, nb < ns ? nb : 0, kl, E01, E23, c_r[j], c_r[j + 1], ta, tb, kq_r[j], kq_r[j + 1], eta,
                                          g01, g23);
                    tau_r[j] = ta;
                    tau_r[j + 1] = tb;
                } else if (ra < trounds) {
                    tauo_r[j] = tau_r[j];
                    tau_r[j] = flda_token1<LPT, CPL>(tile, RS, CH, na < ns ? na : 0, kl, E01, E23, c_r[j], tau_r[j], kq_r[j], eta, g01, g23);
                }
            }
            for (int r = trounds; r < rounds; r++) {   // beyond the tile: rows from L2, tau in global memory (flda_estep_kernel's round)
                if (r % W != warp) continue;
                const int n = r * S + ts;
                const bool ok = n < Nd;
                const int nn = ok ? n : Nd - 1;
                const int term = __ldg(p.terms + o + nn);
                const float kq = __ldg(p.kq + term);
                ulonglong2 b[CPL];
                flda_load_row<LPT, CPL>(tile, p.L, RS, K_ld, CH, nn, 0, term, kl, b);
                const float c = ok ? __ldg(p.counts + o + nn) : 0.0f;
                const float tau = __ldcg(p.tau + o + nn);
                f32x2 p01[CPL], p23[CPL];
                float s, q;
                flda_row<CPL>(b, E01, E23, tau, p01, p23, s, q);
                s = group_sum<LPT>(s);
                q = group_sum<LPT>(q);
                const float rs = rcp_ftz(s);
                const float t = c * rs;
                const f32x2 t2 = pk2(t, t);
#pragma unroll
                for (int m = 0; m < CPL; m++) {
                    g01[m] = fma2(p01[m], t2, g01[m]);
                    g23[m] = fma2(p23[m], t2, g23[m]);
                }
                __syncwarp();   // every lane of the token (and the empty slots that alias the last token) has read tau
                if (ok && kl == 0) {
                    __stcg(p.tau_old + o + nn, tau);
                    __stcg(p.tau + o + nn, flda_new_tau(eta, kq, q, rs));
                }
            }
#pragma unroll
            for (int m = 0; m < CPL; m++)
                if (kl + LPT * m < CH) {
                    float4 gv;
                    unpk2(g01[m], gv.x, gv.y);
                    unpk2(g23[m], gv.z, gv.w);
                    reinterpret_cast<float4 *>(gs_w + (size_t)ts * RS)[kl + LPT * m] = gv;
                }
            flda_cta_sync<W>();
            // ---- K phase: update_gamma! (fLDA.jl:182-185), update_Elogtheta! (fLDA.jl:175-178) for this warp's topic slices
            float dpart = 0.0f, mxn = -3.0e38f;
            v++;
            float *Enn = En_s + (v & 1) * RS;   // last read two token phases ago
#pragma unroll
            for (int j = 0; j < RW; j++) {
                const int i = lane + 32 * (warp + W * j);
                if (i < K) {
                    float gi = owner_sum<S>(gs, RS, i);
#pragma unroll
                    for (int w = 1; w < W; w++) gi += owner_sum<S>(gs + (size_t)w * S * RS, RS, i);
                    gam_k[j] = (alpha_k[j] + gi) + TMVB_EPS;
                    Eold_k[j] = Enew_k[j];
                    Enew_k[j] = psi_lgamma<false, true>(gam_k[j]).psi - psi_sum;
                    const float df = Enew_k[j] - Eold_k[j];
                    dpart = fmaf(df, df, dpart);
                    mxn = fmaxf(mxn, Enew_k[j]);
                    Enn[i] = Enew_k[j];
                }
            }
            dpart = warp_sum(dpart);
            mxn = warp_max(mxn);
            if (W > 1) {
                if (lane == 0) {
                    dpart_s[warp] = dpart;
                    mx_s[warp] = mxn;
                }
                __syncthreads();
                dpart = dpart_s[0];
                mxn = mx_s[0];
#pragma unroll
                for (int w = 1; w < W; w++) {
                    dpart += dpart_s[w];
                    mxn = fmaxf(mxn, mx_s[w]);
                }
            } else {
                __syncwarp();   // the owner sums have been read before gs is overwritten; the new Elogtheta is visible
            }
            // fLDA.jl:229: stop when ||Elogtheta - Elogtheta_old||_2 < vtol (or after viter sweeps); buffer (v - 1) & 1 and mx then
            // still are what the last phi was computed from
            if (v >= p.viter || sqrtf(dpart) < p.vtol) break;
            mx = mxn;
        }

        // ---- update_beta!(model, d) (fLDA.jl:168-171), update_kappa!(model, d) (fLDA.jl:156-159): the last phi, rebuilt from the
        // tau it was computed from (tau_old) and the Elogtheta of the last sweep; weights tau_n c_n and (1 - tau_n) c_n with the FINAL tau
        float tc = 0.0f;
        if (!(p.dbg & 2)) {
            const float *En = En_s + ((v - 1) & 1) * RS;
            f32x2 E01[CPL], E23[CPL];
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                const bool in = (kl + LPT * m < CH);
                const float4 E = in ? reinterpret_cast<const float4 *>(En)[kl + LPT * m] : make_float4(kPadRaw, kPadRaw, kPadRaw, kPadRaw);
                E01[m] = pk2((E.x - mx) * kLog2e, (E.y - mx) * kLog2e);
                E23[m] = pk2((E.z - mx) * kLog2e, (E.w - mx) * kLog2e);
            }
            auto scatter = [&](const ulonglong2(&b)[CPL], bool ok, int term, float c, float tauo, float tauf) {
                f32x2 p01[CPL], p23[CPL];
                float s, q;
                flda_row<CPL>(b, E01, E23, tauo, p01, p23, s, q);
                s = group_sum<LPT>(s);
                if (ok) {
                    const float w = tauf * c * rcp_ftz(s);
                    const f32x2 w2 = pk2(w, w);
                    float *srow = p.stats + (size_t)term * K_ld + 4 * kl;
                    if (!(p.dbg & 1)) {
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
                    if (kl == 0) tc = fmaf(tauf, c, tc);
                }
            };
#pragma unroll
            for (int j = 0; j < TR; j++) {
                if (warp + W * j < trounds) {
                    const int n = (warp + W * j) * S + ts;
                    const bool ok = n < ns;
                    ulonglong2 b[CPL];
                    flda_tile_row<LPT, CPL>(tile, RS, CH, ok ? n : 0, kl, b);
                    scatter(b, ok, term_r[j], c_r[j], tauo_r[j], tau_r[j]);
                    if (ok && kl == 0) {
                        p.tau[o + n] = tau_r[j];
                        p.tau_old[o + n] = tauo_r[j];
                    }
                }
            }
            for (int r = trounds; r < rounds; r++) {
                if (r % W != warp) continue;
                const int n = r * S + ts;
                const bool ok = n < Nd;
                const int nn = ok ? n : Nd - 1;
                const int term = __ldg(p.terms + o + nn);
                ulonglong2 b[CPL];
                flda_load_row<LPT, CPL>(tile, p.L, RS, K_ld, CH, nn, 0, term, kl, b);
                scatter(b, ok, term, __ldg(p.counts + o + nn), __ldcg(p.tau_old + o + nn), __ldcg(p.tau + o + nn));
            }
        }
        tc = warp_sum(tc);
#pragma unroll
        for (int j = 0; j < RW; j++) {
            const int i = lane + 32 * (warp + W * j);
            if (i < K_ld) {
                const bool ok = i < K;
                p.gamma[(size_t)d * K_ld + i] = ok ? gam_k[j] : 0.0f;
                p.Elogtheta[(size_t)d * K_ld + i] = ok ? Enew_k[j] : 0.0f;
                p.Elogtheta_old[(size_t)d * K_ld + i] = ok ? Eold_k[j] : 0.0f;
                if (ok) esum_k[j] += (double)Enew_k[j];
            }
        }
        if (W > 1 && lane == 0) tc_s[warp] = tc;
        flda_cta_sync<W>();   // the tile, term_s, gs and En_s are free for the next document
        if (tid == 0) {
            if (W > 1) {
                tc = tc_s[0];
#pragma unroll
                for (int w = 1; w < W; w++) tc += tc_s[w];
            }
            p.doc_tc[d] = tc;
            tc_thr += (double)tc;
            sweeps_thr += (unsigned long long)v;
        }
    }
#pragma unroll
    for (int j = 0; j < RW; j++) {
        const int i = lane + 32 * (warp + W * j);
        if (i < K && esum_k[j] != 0.0) atomicAdd(p.small + i, esum_k[j]);
    }
    if (tid == 0) {
        if (sweeps_thr) atomicAdd(p.small + K_ld, (double)sweeps_thr);
        if (tc_thr != 0.0) atomicAdd(p.small + K_ld + 1, tc_thr);
    }
}

// update_elbo! (fLDA.jl:62-117), literally: phi rebuilt from (tau_old, beta_old, Elogtheta_old) (fLDA.jl:108), every other
// quantity current.  L_old / L_new are the log2 tables of beta_old / beta; per (token, topic) one FFMA2 + MUFU.EX2 for phi and
// two FFMA2 for  phi (Elogtheta_i + tau ln(beta_i + eps) - ln phi_i),  ln phi_i = ln 2 (x_i - log2 s).
template <int LPT, int CPL>
__global__ void __launch_bounds__(128) flda_elbo_kernel(const FldaDev p, const float *__restrict__ L_old, const float *__restrict__ kappa, double lg_alpha_term,
                                                        double ln_eta, double ln_1m_eta, double *out)
{
    constexpr int S = 32 / LPT;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int kl = lane % LPT, ts = lane / LPT;
    const int K = p.K, K_ld = p.K_ld, CH = K_ld >> 2;
    const ulonglong2 zero = make_ulonglong2(0ull, 0ull);
    double acc = 0.0;
    for (long long d = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); d < p.M; d += (long long)gridDim.x * wpb) {
        const long long o = p.doc_off[d];
        const int Nd = (int)(p.doc_off[d + 1] - o);
        const float *En = p.Elogtheta + d * K_ld, *Eo = p.Elogtheta_old + d * K_ld, *gm = p.gamma + d * K_ld;
        double dacc = 0.0, g0 = 0.0;
        float mx = -3.0e38f;
        for (int i = lane; i < K; i += 32) {
            const double g = gm[i];
            g0 += g;
            const PsiLg pl = psi_lgamma<true>((float)g);
            // Elogptheta + the per-topic part of the Dirichlet entropy (utils.jl:163-180: zero for K = 1)
            dacc += ((double)p.alpha[i] - 1.0) * (double)En[i] + (K > 1 ? (double)pl.lg - (g - 1.0) * (double)pl.psi : 0.0);
            mx = fmaxf(mx, Eo[i]);
        }
        g0 = warp_sum_d(g0);
        mx = warp_max(mx);
        f32x2 Eo01[CPL], Eo23[CPL], En01[CPL], En23[CPL];
#pragma unroll
        for (int m = 0; m < CPL; m++) {
            const int q = kl + LPT * m;
            const bool in = q < CH;
            float4 a = make_float4(kPadE, kPadE, kPadE, kPadE), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in) {
                const float4 eo = *reinterpret_cast<const float4 *>(Eo + 4 * q);
                b4 = *reinterpret_cast<const float4 *>(En + 4 * q);
                a.x = (4 * q + 0 < K) ? (eo.x - mx) * kLog2e : kPadE;
                a.y = (4 * q + 1 < K) ? (eo.y - mx) * kLog2e : kPadE;
                a.z = (4 * q + 2 < K) ? (eo.z - mx) * kLog2e : kPadE;
                a.w = (4 * q + 3 < K) ? (eo.w - mx) * kLog2e : kPadE;
            }
            Eo01[m] = pk2(a.x, a.y);
            Eo23[m] = pk2(a.z, a.w);
            En01[m] = pk2(b4.x, b4.y);
            En23[m] = pk2(b4.z, b4.w);
        }
        float tacc = 0.0f, tc = 0.0f;
        const int rounds = (Nd + S - 1) / S;
        const f32x2 mln2 = pk2(-kLn2, -kLn2);
#pragma unroll 2
        for (int r = 0; r < rounds; r++) {
            const int n = r * S + ts;
            const bool ok = n < Nd;
            const int nn = ok ? n : 0;
            const int term = __ldg(p.terms + o + nn);
            const float c = ok ? __ldg(p.counts + o + nn) : 0.0f;
            const float tau = p.tau[o + nn], tauo = p.tau_old[o + nn];
            const ulonglong2 *ro = reinterpret_cast<const ulonglong2 *>(L_old + (size_t)term * K_ld) + kl;
            const ulonglong2 *rn = reinterpret_cast<const ulonglong2 *>(p.L + (size_t)term * K_ld) + kl;
            const f32x2 to2 = pk2(tauo, tauo), tl2 = pk2(tau * kLn2, tau * kLn2);
            f32x2 sa = 0ull, sb = 0ull, Aa = 0ull, Ab = 0ull;
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                const bool in = (kl + LPT * m < CH);
                const ulonglong2 bo = in ? __ldg(ro + LPT * m) : zero, bn = in ? __ldg(rn + LPT * m) : zero;
                const f32x2 xa = fma2(to2, bo.x, Eo01[m]), xb = fma2(to2, bo.y, Eo23[m]);
                float x0, x1, x2, x3;
                unpk2(xa, x0, x1);
                unpk2(xb, x2, x3);
                // pad topics: x = kPadE, p = 0, and p * (finite) = 0
                const f32x2 pa = pk2(ex2_ftz(x0), ex2_ftz(x1)), pb = pk2(ex2_ftz(x2), ex2_ftz(x3));
                sa = add2(sa, pa);
                sb = add2(sb, pb);
                Aa = fma2(pa, fma2(mln2, xa, fma2(tl2, bn.x, En01[m])), Aa);
                Ab = fma2(pb, fma2(mln2, xb, fma2(tl2, bn.y, En23[m])), Ab);
            }
            const float s = group_sum<LPT>(hsum2(add2(sa, sb)));
            const float A = group_sum<LPT>(hsum2(add2(Aa, Ab)));
            if (ok && kl == 0) {
                // Elogpz + Elogpw (topical part) - Elogqz of the token, then the corpus part of Elogpw and -Elogqc
                float tok = __fdividef(A, s) + __logf(s);
                tok = fmaf(1.0f - tau, logf(__ldg(kappa + term) + TMVB_EPS), tok);
                const float t0 = 1.0f - tau;
                if (t0 != 0.0f && t0 != 1.0f) tok -= t0 * logf(t0) + tau * logf(tau);
                tacc = fmaf(c, tok, tacc);
                tc = fmaf(tau, c, tc);
            }
        }
        dacc += (double)tacc;
        dacc = warp_sum_d(dacc);
        tc = warp_sum(tc);
        if (lane == 0) {
            double ent = 0.0;
            if (K > 1) ent = -lgamma(g0) + (g0 - (double)K) * (double)psi_lgamma<false>((float)g0).psi;
            // Elogpc (fLDA.jl:68-72): log(@boink eta^(tau . c) (1 - eta)^(C_d - tau . c))
            const double Cd = (double)p.doc_c[d];
            const double elogpc = log(exp((double)tc * ln_eta + (Cd - (double)tc) * ln_1m_eta) + TMVB_EPS_D);
            acc += dacc + ent + lg_alpha_term + elogpc;
        }
    }
    if (lane == 0 && acc != 0.0) atomicAdd(out, acc);
}

// L[j][i] = log2(beta[j][i] + eps), pad topics 0
__global__ void flda_logtable_kernel(const float *__restrict__ beta, float *__restrict__ L, long long n, int K, int K_ld)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q % K_ld);
        L[q] = (i < K) ? log2f(beta[q] + TMVB_EPS) : 0.0f;
    }
}
// update_kappa!(model) (fLDA.jl:149-153): kappa_old <- kappa; kappa = kappa_temp ./ sum(kappa_temp); kappa_temp <- 0.  One CTA.
__global__ void flda_kappa_kernel(float *__restrict__ kstats, float *__restrict__ kappa, float *__restrict__ kappa_old, int V)
{
    __shared__ double red[32];
    __shared__ double tot;
    double a = 0.0;
    for (int j = threadIdx.x; j < V; j += blockDim.x) a += (double)kstats[j];
    a = warp_sum_d(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
        tot = t;
    }
    __syncthreads();
    const double inv = 1.0 / tot;
    for (int j = threadIdx.x; j < V; j += blockDim.x) {
        kappa_old[j] = kappa[j];
        kappa[j] = (float)((double)kstats[j] * inv);
        kstats[j] = 0.0f;
    }
}
__global__ void flda_kq_kernel(const float *__restrict__ kappa, float *__restrict__ kq, int V, float eta)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < V; j += gridDim.x * blockDim.x) kq[j] = (1.0f - eta) * kappa[j];
    if (blockIdx.x == 0 && threadIdx.x == 0) kq[V] = eta;
}
// per-token arrays between the caller's CSR order and the shard's internal order (documents sorted by length)
__global__ void flda_tok_permute_kernel(const float *__restrict__ src, float *__restrict__ dst, const long long *__restrict__ src_off,
                                        const long long *__restrict__ dst_off, long long M, int to_internal, int *__restrict__ err)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    int e = 0;
    for (long long d = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); d < M; d += (long long)gridDim.x * wpb) {
        const long long so = src_off[d], o = dst_off[d];
        const int Nd = (int)(dst_off[d + 1] - o);
        for (int n = lane; n < Nd; n += 32) {
            if (to_internal) {
                const float v = src[so + n];
                if (!(v >= 0.0f && v <= 1.0f)) e = 1;
                dst[o + n] = v;
            } else {
                dst[so + n] = src[o + n];
            }
        }
    }
    if (e && err) atomicOr(err, 1);
}

// ---- host-side launches shared with the filtered CTM (tmvb_filt.cuh) --------------------------------------------------
int filt_log_table(Shard *s, const float *beta, float *L)
{
    const long long n = (long long)s->V * s->K_ld;
    if (n == 0) return 0;
    flda_logtable_kernel<<<grid_for(n, 256, s->n_sm), 256, 0, s->stream>>>(beta, L, n, (int)s->K, s->K_ld);
    TMVB_CUDA(cudaGetLastError());
    s->st.kernel_launches++;
    return 0;
}
int filt_kappa_update(Shard *s, float *kstats, float *kappa, float *kappa_old)
{
    if (s->V == 0) return 0;
    flda_kappa_kernel<<<1, 1024, 0, s->stream>>>(kstats, kappa, kappa_old, (int)s->V);
    TMVB_CUDA(cudaGetLastError());
    s->st.kernel_launches++;
    return 0;
}
int filt_push_kq(Shard *s, const float *kappa, float *kq, double eta)
{
    flda_kq_kernel<<<grid_for(std::max<int64_t>(s->V, 1), 256, s->n_sm), 256, 0, s->stream>>>(kappa, kq, (int)s->V, (float)eta);
    TMVB_CUDA(cudaGetLastError());
    s->st.kernel_launches++;
    return 0;
}
int filt_tau_upload(Shard *s, const float *host_tau, float *d_tau, int *bad)
{
    *bad = 0;
    if (s->nnz == 0) return 0;
    TMVB_TRY(shard_scratch(s, (size_t)s->nnz * 4));
    TMVB_CUDA(cudaMemcpyAsync(s->d_scratch, host_tau, (size_t)s->nnz * 4, cudaMemcpyHostToDevice, s->stream));
    TMVB_CUDA(cudaMemsetAsync(s->d_counters + 61, 0, 4, s->stream));
    flda_tok_permute_kernel<<<grid_for(s->M * 32, 256, s->n_sm), 256, 0, s->stream>>>((const float *)s->d_scratch, d_tau, s->d_src_off, s->d_doc_off, s->M, 1,
                                                                                 s->d_counters + 61);
    TMVB_CUDA(cudaGetLastError());
    TMVB_CUDA(cudaMemcpyAsync(bad, s->d_counters + 61, 4, cudaMemcpyDeviceToHost, s->stream));
    TMVB_CUDA(cudaStreamSynchronize(s->stream));
    s->st.h2d_bytes += s->nnz * 4;
    s->st.kernel_launches++;
    return 0;
}
int filt_tau_download(Shard *s, const float *d_tau, float *host_tau)
{
    if (s->nnz == 0) return 0;
    TMVB_TRY(shard_scratch(s, (size_t)s->nnz * 4));
    flda_tok_permute_kernel<<<grid_for(s->M * 32, 256, s->n_sm), 256, 0, s->stream>>>(d_tau, (float *)s->d_scratch, s->d_src_off, s->d_doc_off, s->M, 0, nullptr);
    TMVB_CUDA(cudaGetLastError());
    TMVB_CUDA(cudaMemcpyAsync(host_tau, s->d_scratch, (size_t)s->nnz * 4, cudaMemcpyDeviceToHost, s->stream));
    TMVB_CUDA(cudaStreamSynchronize(s->stream));
    s->st.d2h_bytes += s->nnz * 4;
    s->st.kernel_launches++;
    return 0;
}

typedef void (*FldaEstepFn)(const FldaDev, int, int, int, int, int *);
typedef void (*FldaElboFn)(const FldaDev, const float *, const float *, double, double, double, double *);
#define TMVB_FLDA_FN(L, C) (FldaEstepFn)flda_estep_kernel<L, C>,
#define TMVB_FLDA_ELBO_FN(L, C) (FldaElboFn)flda_elbo_kernel<L, C>,
static const FldaEstepFn kFldaEstep[kNumLaneLayouts] = {TMVB_FOR_EACH_LAYOUT(TMVB_FLDA_FN)};
// the register-state variant exists for the two-lanes-per-token layouts (K_ld <= 64: 16 token streams, 4 tile rounds per warp);
// [0]: one warp per document (tiles up to 64 tokens), [1]: two warps per document (tiles up to 128 tokens)
constexpr int kFldaRegTile = 64;   // tokens one warp keeps register state for
template <int L, int C, int W, int MAXREG>
constexpr FldaEstepFn flda_reg_fn()
{
    if constexpr (L == 2)
        return (FldaEstepFn)flda_estep_reg_kernel<L, C, kFldaRegTile / (32 / L), W, MAXREG>;
    else
        return nullptr;
}
// [layout][warps per document - 1]; capped at 168 registers per thread (12 resident warps per SM; a dozen spilled words outside
// the token loop): NSF K=50 E-step 3.55 ms, against 3.76 ms at 200 registers (10 warps, no spills)
#define TMVB_FLDA_REG_FN(L, C) {flda_reg_fn<L, C, 1, 168>(), flda_reg_fn<L, C, 2, 168>()},
static const FldaEstepFn kFldaEstepReg[kNumLaneLayouts][2] = {TMVB_FOR_EACH_LAYOUT(TMVB_FLDA_REG_FN)};

struct FldaPick {
    const void *tile, *reg[2];
};
const void *flda_pick(const Bucket &b, const void *ctx)
{
    const FldaPick *pk = static_cast<const FldaPick *>(ctx);
    return b.hyb ? pk->reg[b.warps - 1] : pk->tile;
}
// TMVB_FLDA_REG (default 2): 0 = flda_estep_kernel everywhere (64-token tiles), 1 = register-state variant with one warp per document
// (64-token tiles), 2 = ... and two warps per document on the launches with 80- to 128-token tiles
int flda_reg_mode(const Shard &s)
{
    const int v = env_int("TMVB_FLDA_REG", 2);
    return (v >= 1 && v <= 2 && kFldaEstepReg[s.layout][0]) ? v : 0;
}
static const FldaElboFn kFldaElbo[kNumLaneLayouts] = {TMVB_FOR_EACH_LAYOUT(TMVB_FLDA_ELBO_FN)};

}  // namespace tmvb

using namespace tmvb;

struct tmvb_flda_s {
    Shard s;
    bool params_set = false, no_scatter = false;
    double eta = 0.5;
    float *d_alpha = nullptr;
    double *d_alpha64 = nullptr;
    float *d_L[2] = {nullptr, nullptr};   // log2 tables of s.d_beta[0 / 1]
    float *d_kappa = nullptr, *d_kappa_old = nullptr, *d_kstats = nullptr, *d_kq = nullptr;
    float *d_Elogtheta = nullptr, *d_Elogtheta_old = nullptr, *d_gamma = nullptr;
    float *d_tau = nullptr, *d_tau_old = nullptr, *d_doc_tc = nullptr;
    size_t tau_cap = 0;
    int reg_mode = 0;            // TMVB_FLDA_REG at create: 0 tile kernel, 1 / 2 register-state kernel with up to 1 / 2 warps per document
    double *d_small = nullptr;   // [K_ld] sum_d Elogtheta | sweeps | sum tau c | ELBO
    double *d_local = nullptr;   // [2 K_ld] rowsum | elbo_w (shard_normalize)
    int64_t n_small = 0;
    Comm comm;   // peer-memory all-reduce of the statistics (multi-GPU)
};

namespace {

FldaDev flda_view(tmvb_flda_t h)
{
    Shard &s = h->s;
    FldaDev p;
    memset(&p, 0, sizeof(p));
    p.K = (int)s.K;
    p.K_ld = s.K_ld;
    p.V = (int)s.V;
    p.RS = s.RS;
    p.M = s.M;
    p.L = h->d_L[s.cur];
    p.kq = h->d_kq;
    p.alpha = h->d_alpha;
    p.stats = s.d_stats;
    p.kstats = h->d_kstats;
    p.doc_off = s.d_doc_off;
    p.terms = s.d_terms;
    p.counts = s.d_counts;
    p.doc_c = s.d_doc_c;
    p.Elogtheta = h->d_Elogtheta;
    p.Elogtheta_old = h->d_Elogtheta_old;
    p.gamma = h->d_gamma;
    p.tau = h->d_tau;
    p.tau_old = h->d_tau_old;
    p.doc_tc = h->d_doc_tc;
    p.small = h->d_small;
    p.dbg = env_int("TMVB_DBG", 0) | (h->no_scatter ? 1 : 0);
    return p;
}

void flda_free(tmvb_flda_t h)
{
    cudaSetDevice(h->s.device);
    if (h->s.stream) cudaStreamSynchronize(h->s.stream);
    cudaFree(h->d_alpha);
    cudaFree(h->d_alpha64);
    cudaFree(h->d_L[0]);
    cudaFree(h->d_L[1]);
    cudaFree(h->d_kappa);
    cudaFree(h->d_kappa_old);
    cudaFree(h->d_kstats);
    cudaFree(h->d_kq);
    cudaFree(h->d_Elogtheta);
    cudaFree(h->d_Elogtheta_old);
    cudaFree(h->d_gamma);
    cudaFree(h->d_tau);
    cudaFree(h->d_tau_old);
    cudaFree(h->d_doc_tc);
    cudaFree(h->d_small);
    cudaFree(h->d_local);
    comm_free(&h->comm);
    shard_free(&h->s);
}

PeerReduce flda_peer_bufs(tmvb_flda_t h)
{
    PeerReduce b;
    b.f[0] = h->s.d_stats;
    b.nf[0] = (long long)h->s.V * h->s.K_ld;
    b.f[1] = h->d_kstats;
    b.nf[1] = ((long long)std::max<int64_t>(h->s.V, 1) + 3) / 4 * 4;
    b.small = h->d_small;
    b.n_small = h->n_small;
    return b;
}

int flda_log_table(tmvb_flda_t h, int which) { return filt_log_table(&h->s, h->s.d_beta[which], h->d_L[which]); }

int flda_push_kq(tmvb_flda_t h) { return filt_push_kq(&h->s, h->d_kappa, h->d_kq, h->eta); }

}  // namespace

extern "C" {

int tmvb_flda_create(tmvb_flda_t *out, int64_t K, int64_t M, int64_t V, int device, void *stream)
{
    TMVB_CHECK_ARG(out != nullptr, "handle pointer is NULL");
    *out = nullptr;
    TMVB_CHECK_ARG(K > 0, "number of topics must be a positive integer");  // fLDA.jl:32
    tmvb_flda_t h = new tmvb_flda_s();
    const int64_t K_ld = (K + 7) / 8 * 8;
    h->n_small = K_ld + 3;
    int rc = shard_create(&h->s, K, M, V, device, stream, (size_t)h->n_small + 2 * K_ld + 8);
    if (rc == 0) {
        Shard &s = h->s;
        s.per_tok_extra = kFldaPerTokExtra;
        const size_t km = (size_t)std::max<int64_t>(M, 1) * s.K_ld, kv = (size_t)std::max<int64_t>(V, 1) * s.K_ld, v1 = (size_t)std::max<int64_t>(V, 1);
        cudaError_t e = cudaSuccess;
        auto A = [&](void **p, size_t bytes) {
            if (e == cudaSuccess) e = cudaMalloc(p, bytes);
            if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes, s.stream);
        };
        A((void **)&h->d_alpha, s.K_ld * 4);
        A((void **)&h->d_alpha64, s.K_ld * 8);
        A((void **)&h->d_L[0], kv * 4);
        A((void **)&h->d_L[1], kv * 4);
        A((void **)&h->d_kappa, v1 * 4);
        A((void **)&h->d_kappa_old, v1 * 4);
        A((void **)&h->d_kstats, (v1 + 3) / 4 * 16);   // a multiple of four floats: the peer all-reduce moves 16-byte elements
        A((void **)&h->d_kq, (v1 + 1) * 4);
        A((void **)&h->d_Elogtheta, km * 4);
        A((void **)&h->d_Elogtheta_old, km * 4);
        A((void **)&h->d_gamma, km * 4);
        A((void **)&h->d_doc_tc, (size_t)std::max<int64_t>(M, 1) * 4);
        A((void **)&h->d_small, (h->n_small + 1) * 8);
        A((void **)&h->d_local, (2 * s.K_ld + 2) * 8);
        if (e == cudaSuccess) e = cudaFuncSetAttribute((const void *)kFldaEstep[s.layout], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem_optin);
        for (int v = 0; v < 2; v++)
            if (e == cudaSuccess && kFldaEstepReg[s.layout][v])
                e = cudaFuncSetAttribute((const void *)kFldaEstepReg[s.layout][v], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem_optin);
        // a tile of at most 64 tokens per warp (longer documents read their remaining rows from L2): NSF K=50 E-step 5.23 ms with tiles
        // sized for the longest document of a launch (2-5 resident warps per SM), 4.47 / 4.18 / 3.93 / 3.91 / 4.30 ms at 16 / 32 / 48 / 64 / 96
        h->reg_mode = flda_reg_mode(s);
        s.tile_cap_max = kFldaRegTile * (h->reg_mode == 2 ? 2 : 1);
        if (e != cudaSuccess) rc = fail((int)e, "device allocation failed: %s", cudaGetErrorString(e));
    }
    if (rc != 0) {
        flda_free(h);
        delete h;
        return rc;
    }
    *out = h;
    return 0;
}

int tmvb_flda_destroy(tmvb_flda_t h)
{
    if (!h) return 0;
    flda_free(h);
    delete h;
    return 0;
}

int tmvb_flda_set_corpus(tmvb_flda_t h, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    const size_t fixed = h->reg_mode ? flda_reg_fixed_smem(s.RS, s.lpt, h->reg_mode) : flda_fixed_smem(s.RS, s.lpt);
    TMVB_TRY(shard_set_corpus(&s, N_cumsum, terms, counts, fixed));
    const size_t per_tok = (size_t)s.RS * 4 + 8 + s.per_tok_extra;
    for (Bucket &b : s.buckets) {
        // register-state kernel where the tile is a whole number of 16-token rounds its warps can keep state for
        b.hyb = (h->reg_mode && b.cap % 16 == 0 && b.cap <= kFldaRegTile * h->reg_mode) ? 1 : 0;
        b.warps = (b.hyb && b.cap > kFldaRegTile) ? 2 : 1;
        b.smem = (b.hyb ? flda_reg_fixed_smem(s.RS, s.lpt, b.warps) : flda_fixed_smem(s.RS, s.lpt)) + (size_t)b.cap * per_tok;
        b.grid = 0;
    }
    const size_t need = (size_t)std::max<int64_t>(s.nnz, 1);
    if (need > h->tau_cap) {
        TMVB_CUDA(cudaSetDevice(s.device));
        cudaFree(h->d_tau);
        cudaFree(h->d_tau_old);
        h->d_tau = h->d_tau_old = nullptr;
        h->tau_cap = 0;
        TMVB_CUDA(cudaMalloc((void **)&h->d_tau, need * 4));
        TMVB_CUDA(cudaMalloc((void **)&h->d_tau_old, need * 4));
        h->tau_cap = need;
    }
    return 0;
}

/* eta, alpha[K], kappa[V], beta[K*V], Elogtheta[K*M], gamma[K*M], tau[nnz] (the caller's CSR order); any pointer may be NULL */
int tmvb_flda_upload(tmvb_flda_t h, const double *eta, const float *alpha, const float *kappa, const float *beta, const float *Elogtheta,
                     const float *gamma, const float *tau)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    const int K = (int)s.K;
    if (eta) {
        if (!(*eta >= 0.0 && *eta <= 1.0)) return fail(-5, "eta must belong to the interval [0,1].");   // modelutils.jl:71
        h->eta = *eta;
    }
    if (alpha) {
        std::vector<float> a32(s.K_ld, 0.f);
        std::vector<double> a64(s.K_ld, 0.0);
        for (int i = 0; i < K; i++) {
            if (!isfinite(alpha[i])) return fail(-5, "alpha must be finite.");
            if (!(alpha[i] > 0.f)) return fail(-5, "alpha must be positive.");
            a32[i] = alpha[i];
            a64[i] = (double)alpha[i];
        }
        TMVB_CUDA(cudaMemcpyAsync(h->d_alpha, a32.data(), s.K_ld * 4, cudaMemcpyHostToDevice, s.stream));
        TMVB_CUDA(cudaMemcpyAsync(h->d_alpha64, a64.data(), s.K_ld * 8, cudaMemcpyHostToDevice, s.stream));
        TMVB_CUDA(cudaStreamSynchronize(s.stream));
        s.st.h2d_bytes += s.K * 12;
    }
    if (kappa && s.V > 0) {
        double ks = 0.0;
        for (int64_t j = 0; j < s.V; j++) {
            if (!(kappa[j] >= 0.f) || !isfinite(kappa[j])) return fail(-5, "kappa must be a probability vector of length V.");   // modelutils.jl:74
            ks += kappa[j];
        }
        if (fabs(ks - 1.0) > 1e-3) return fail(-5, "kappa must be a probability vector of length V.");
        TMVB_CUDA(cudaMemcpyAsync(h->d_kappa, kappa, s.V * 4, cudaMemcpyHostToDevice, s.stream));
        TMVB_CUDA(cudaMemcpyAsync(h->d_kappa_old, h->d_kappa, s.V * 4, cudaMemcpyDeviceToDevice, s.stream));   // kappa_old = copy(kappa), fLDA.jl:42
        TMVB_CUDA(cudaStreamSynchronize(s.stream));
        s.st.h2d_bytes += s.V * 4;
    }
    TMVB_TRY(flda_push_kq(h));
    if (beta && s.V > 0) {
        TMVB_TRY(shard_upload_rows(&s, beta, s.d_beta[s.cur], s.V, nullptr, 0));
        TMVB_TRY(shard_check_stochastic(&s, s.d_beta[s.cur]));   // the row sums of check_model, on the device copy
        TMVB_CUDA(cudaMemcpyAsync(s.d_beta[s.cur ^ 1], s.d_beta[s.cur], (size_t)s.V * s.K_ld * 4, cudaMemcpyDeviceToDevice, s.stream));   // fLDA.jl:45
        TMVB_TRY(flda_log_table(h, 0));
        TMVB_TRY(flda_log_table(h, 1));
    }
    if ((Elogtheta || gamma || tau) && s.M > 0) {
        TMVB_CHECK_ARG(s.corpus_set, "set_corpus must precede the upload of per-document parameters");
        TMVB_TRY(shard_upload_rows(&s, Elogtheta, h->d_Elogtheta, s.M, s.d_perm, 1));
        if (Elogtheta)   // Elogtheta_old = deepcopy(Elogtheta), fLDA.jl:48
            TMVB_CUDA(cudaMemcpyAsync(h->d_Elogtheta_old, h->d_Elogtheta, (size_t)s.M * s.K_ld * 4, cudaMemcpyDeviceToDevice, s.stream));
        TMVB_TRY(shard_upload_rows(&s, gamma, h->d_gamma, s.M, s.d_perm, 2));
        if (tau && s.nnz > 0) {
            int terr = 0;
            TMVB_TRY(filt_tau_upload(&s, tau, h->d_tau, &terr));
            TMVB_CUDA(cudaMemcpyAsync(h->d_tau_old, h->d_tau, (size_t)s.nnz * 4, cudaMemcpyDeviceToDevice, s.stream));   // tau_old = deepcopy(tau), fLDA.jl:51
            if (terr) return fail(-5, "tau must contain probabilities.");
        }
    }
    int verr = 0;
    TMVB_TRY(shard_validation(&s, &verr));
    if (verr & 0x4003) return fail(-5, "beta must be a right stochastic matrix.");   // the messages of check_model(::fLDA), modelutils.jl:69-98
    if (verr & 0x4) return fail(-5, "Elogtheta must be finite.");
    if (verr & 0x8) return fail(-5, "Elogtheta must be nonpositive.");
    if (verr & 0x10) return fail(-5, "gamma must be finite.");
    if (verr & 0x20) return fail(-5, "gamma must be positive.");
    h->params_set = true;
    return 0;
}

/* the inner loop of train!(::fLDA) (fLDA.jl:223-233) for every document of the shard, then update_beta!(model, d) and
 * update_kappa!(model, d) (scatter) */
int tmvb_flda_estep(tmvb_flda_t h, int viter, float vtol)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CHECK_ARG(viter >= 1, "viter must be at least 1");
    TMVB_CHECK_ARG(vtol >= 0.f, "tolerance parameters must be nonnegative");   // fLDA.jl:216
    Shard &s = h->s;
    TMVB_CHECK_ARG(s.corpus_set && h->params_set, "set_corpus / upload have not been called");
    TMVB_CUDA(cudaSetDevice(s.device));
    FldaDev p = flda_view(h);
    p.viter = viter;
    p.vtol = vtol;
    TMVB_CUDA(cudaEventRecord(s.ev[0], s.stream));
    TMVB_CUDA(cudaMemsetAsync(h->d_small, 0, h->n_small * 8, s.stream));
    const FldaPick pk = {(const void *)kFldaEstep[s.layout], {(const void *)kFldaEstepReg[s.layout][0], (const void *)kFldaEstepReg[s.layout][1]}};
    TMVB_TRY(shard_launch(&s, flda_pick, &pk, &p, sizeof(p)));
    TMVB_CUDA(cudaEventRecord(s.ev[1], s.stream));
    s.estep_timed = true;
    return 0;
}

/* predict(corp, train_model::fLDA) (modelutils.jl:857-884): the inner loop without the scatter */
int tmvb_flda_predict(tmvb_flda_t h, int viter, float vtol)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    h->no_scatter = true;
    const int rc = tmvb_flda_estep(h, viter, vtol);
    h->no_scatter = false;
    return rc;
}

/* the three buffers a multi-GPU caller sums over ranks between estep and mstep */
int tmvb_flda_reduce_buffers(tmvb_flda_t h, void **stats, int64_t *n_stats, void **kstats, int64_t *n_kstats, void **small, int64_t *n_small)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    if (stats) *stats = h->s.d_stats;
    if (n_stats) *n_stats = (int64_t)h->s.V * h->s.K_ld;
    if (kstats) *kstats = h->d_kstats;
    if (n_kstats) *n_kstats = h->s.V;
    if (small) *small = h->d_small;
    if (n_small) *n_small = h->n_small;
    return 0;
}

/* ---- multi-GPU: the statistics summed over the ranks by ONE kernel over CUDA-IPC peer memory (tmvb_peer.cu) instead of one NCCL
 * all-reduce per buffer.  Handshake as for gpuLDA: export -> all-gather the blobs over any transport -> connect; then
 * tmvb_flda_peer_reduce(h) between estep and mstep on every rank. ---- */
int tmvb_flda_comm_export(tmvb_flda_t h, void *blob, int64_t blob_bytes)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CHECK_ARG(blob_bytes >= TMVB_COMM_BLOB_BYTES, "blob must hold TMVB_COMM_BLOB_BYTES");
    TMVB_CUDA(cudaSetDevice(h->s.device));
    return peer_export(&h->comm, flda_peer_bufs(h), blob, (size_t)blob_bytes);
}

int tmvb_flda_comm_connect(tmvb_flda_t h, int rank, int world, const void *blobs, int64_t blob_bytes)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CUDA(cudaSetDevice(h->s.device));
    TMVB_CUDA(cudaStreamSynchronize(h->s.stream));
    return comm_connect(&h->comm, rank, world, blobs, (size_t)blob_bytes);
}

int tmvb_flda_peer_reduce(tmvb_flda_t h)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    TMVB_CUDA(cudaSetDevice(h->s.device));
    TMVB_TRY(peer_allreduce(&h->comm, flda_peer_bufs(h), h->s.stream, h->s.n_sm));
    h->s.st.kernel_launches++;
    return 0;
}

/* update_beta!(model), update_kappa!(model), update_alpha!(model, niter, ntol), update_eta!(model) (fLDA.jl:236-239);
 * C_total = sum(model.C) over ALL ranks */
int tmvb_flda_mstep(tmvb_flda_t h, int64_t M_total, double C_total, int niter, double ntol)
{
    TMVB_CHECK_ARG(h != nullptr && M_total > 0 && C_total > 0.0, "bad arguments");
    TMVB_CHECK_ARG(niter >= 0 && ntol >= 0.0, "iteration/tolerance parameters must be nonnegative");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    TMVB_CUDA(cudaEventRecord(s.ev[2], s.stream));
    TMVB_TRY(shard_normalize(&s, h->d_local, false, false));   // beta_old <- beta; beta = beta_temp ./ rowsum; beta_temp <- 0
    TMVB_TRY(flda_log_table(h, s.cur));
    TMVB_TRY(filt_kappa_update(&s, h->d_kstats, h->d_kappa, h->d_kappa_old));
    TMVB_TRY(lda_launch_alpha(h->d_alpha64, h->d_alpha, h->d_small, (int)s.K, s.K_ld, (double)M_total, niter, ntol, s.stream));
    s.st.kernel_launches++;
    TMVB_CUDA(cudaMemcpyAsync(s.h_pinned, h->d_small + s.K_ld + 1, 8, cudaMemcpyDeviceToHost, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    s.st.d2h_bytes += 8;
    h->eta = s.h_pinned[0] / C_total;   // update_eta!, fLDA.jl:119-121
    if (h->comm.connected) {
        int pst = 0;
        TMVB_TRY(peer_status(&h->comm, s.stream, &pst));
    }
    TMVB_TRY(flda_push_kq(h));
    TMVB_CUDA(cudaEventRecord(s.ev[3], s.stream));
    s.mstep_timed = true;
    return 0;
}

/* update_elbo! (fLDA.jl:105-117) of the shard's documents, from the device state */
int tmvb_flda_elbo(tmvb_flda_t h, double *elbo_docs)
{
    TMVB_CHECK_ARG(h && elbo_docs, "NULL argument");
    Shard &s = h->s;
    TMVB_CHECK_ARG(s.corpus_set && h->params_set, "set_corpus / upload have not been called");
    TMVB_CUDA(cudaSetDevice(s.device));
    const int K = (int)s.K;
    double *out = h->d_small + h->n_small;
    TMVB_CUDA(cudaMemsetAsync(out, 0, 8, s.stream));
    if (s.M > 0) {
        TMVB_CUDA(cudaMemcpyAsync(s.h_pinned, h->d_alpha64, K * 8, cudaMemcpyDeviceToHost, s.stream));
        TMVB_CUDA(cudaStreamSynchronize(s.stream));
        s.st.d2h_bytes += K * 8;
        double a0 = 0.0, sl = 0.0;
        for (int i = 0; i < K; i++) {
            a0 += s.h_pinned[i];
            sl += lgamma(s.h_pinned[i]);
        }
        FldaDev p = flda_view(h);
        const int grid = grid_for(s.M * 32, 128, s.n_sm);
        kFldaElbo[s.layout]<<<grid, 128, 0, s.stream>>>(p, h->d_L[s.cur ^ 1], h->d_kappa, lgamma(a0) - sl, log(h->eta), log1p(-h->eta), out);
        TMVB_CUDA(cudaGetLastError());
        s.st.kernel_launches++;
    }
    TMVB_CUDA(cudaMemcpyAsync(s.h_pinned, out, 8, cudaMemcpyDeviceToHost, s.stream));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    s.st.d2h_bytes += 8;
    *elbo_docs = s.h_pinned[0];
    return 0;
}

int tmvb_flda_download(tmvb_flda_t h, double *eta, float *alpha, float *kappa, float *beta, float *Elogtheta, float *gamma, float *tau)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    if (eta) *eta = h->eta;
    if (alpha) {
        TMVB_CUDA(cudaMemcpyAsync(alpha, h->d_alpha, s.K * 4, cudaMemcpyDeviceToHost, s.stream));
        s.st.d2h_bytes += s.K * 4;
    }
    if (kappa && s.V > 0) {
        TMVB_CUDA(cudaMemcpyAsync(kappa, h->d_kappa, s.V * 4, cudaMemcpyDeviceToHost, s.stream));
        s.st.d2h_bytes += s.V * 4;
    }
    if (beta) TMVB_TRY(shard_download_rows(&s, s.d_beta[s.cur], beta, s.V, nullptr));
    if (Elogtheta) TMVB_TRY(shard_download_rows(&s, h->d_Elogtheta, Elogtheta, s.M, s.d_perm));
    if (gamma) TMVB_TRY(shard_download_rows(&s, h->d_gamma, gamma, s.M, s.d_perm));
    if (tau) TMVB_TRY(filt_tau_download(&s, h->d_tau, tau));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    return 0;
}

int tmvb_flda_download_old(tmvb_flda_t h, float *kappa_old, float *beta_old, float *Elogtheta_old, float *tau_old)
{
    TMVB_CHECK_ARG(h != nullptr, "handle is NULL");
    Shard &s = h->s;
    TMVB_CUDA(cudaSetDevice(s.device));
    if (kappa_old && s.V > 0) {
        TMVB_CUDA(cudaMemcpyAsync(kappa_old, h->d_kappa_old, s.V * 4, cudaMemcpyDeviceToHost, s.stream));
        s.st.d2h_bytes += s.V * 4;
    }
    if (beta_old) TMVB_TRY(shard_download_rows(&s, s.d_beta[s.cur ^ 1], beta_old, s.V, nullptr));
    if (Elogtheta_old) TMVB_TRY(shard_download_rows(&s, h->d_Elogtheta_old, Elogtheta_old, s.M, s.d_perm));
    if (tau_old) TMVB_TRY(filt_tau_download(&s, h->d_tau_old, tau_old));
    TMVB_CUDA(cudaStreamSynchronize(s.stream));
    return 0;
}

int tmvb_flda_topics(tmvb_flda_t h, int32_t *topics) /* fLDA.jl:246 */
{
    TMVB_CHECK_ARG(h && topics, "NULL argument");
    return shard_topics(&h->s, h->s.d_beta[h->s.cur], nullptr, topics);
}

int tmvb_flda_get_stats(tmvb_flda_t h, tmvb_stats *out)
{
    TMVB_CHECK_ARG(h && out, "NULL argument");
    return shard_get_stats(&h->s, h->d_small + h->s.K_ld, out);
}

}  // extern "C"
