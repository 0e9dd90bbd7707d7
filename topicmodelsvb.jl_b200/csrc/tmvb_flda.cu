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
    TMVB_TRY(shard_set_corpus(&s, N_cumsum, terms, counts, flda_fixed_smem(s.RS, s.lpt)));
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
    const void *fns[1] = {(const void *)kFldaEstep[s.layout]};
    TMVB_TRY(shard_launch(&s, pick_by_warps, fns, &p, sizeof(p)));
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
