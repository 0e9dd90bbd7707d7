// tmvb_lda_estep.cuh -- the LDA E-step kernels (tile, register-resident, hybrid), shared by tmvb_lda.cu and by developer
// probes that instantiate single variants (tools/).  See tmvb_lda.cu for the data layout.
#pragma once

#include "tmvb_comm.cuh"
#include "tmvb_shard.cuh"

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
    const float *doc_c;   // [M] sum of a document's counts
    float *Elogtheta, *Elogtheta_old, *gamma;
    double *small;
    int viter;
    float vtol;
    int stage_bulk;  // 1: TMA bulk row copies (UBLKCP), 0: 16-byte cp.async (LDGSTS)
    int dbg;         // developer probes: bit0 skip the scatter, bit1 skip the final pass
    // host mirror (tmvb_lda_arm_host_mirror): when set, the E-step also writes each document's final gamma / Elogtheta row into
    // the caller's page-locked K x M arrays (row perm[d], K floats) -- update_host! overlapped with the sweeps of the other documents
    float *host_E, *host_gamma;
    const int *perm;  // sorted position -> the caller's document index
};

// shared memory of one E-step CTA beyond the tile: 256-byte header (mbarrier, next-document slot, per-warp partial sums) |
// gs [W][S][RS] | e_s [RS]
static size_t lda_fixed_smem(int RS, int lpt, int W) { return 256 + (size_t)W * (32 / lpt) * RS * 4 + (size_t)RS * 4; }

// documents drawn from a bucket's work counter per atomic
constexpr int kDocChunk = 8;

template <int W>
__device__ __forceinline__ void cta_sync()
{
    if (W == 1)
        __syncwarp();
    else
        __syncthreads();
}

// W warps cooperate on one document (W = 1 for short documents, 2 for the rest): they share the staged tile, split the
// token rounds (warp w takes rounds w, w+W, ...) and split the topics of the K phase (thread t owns topics t + 32W r).
// Two CTA barriers per sweep: after the per-stream partial K-vectors are in shared memory, and after exp(Elogtheta) and
// the partial convergence sums are.
template <int LPT, int CPL, int W, bool ELBO>
__global__ void __launch_bounds__(32 * W, W == 2 ? 8 : 1) lda_estep_kernel(const LdaDev p, int doc_begin, int doc_end, int cap, int cap2, int *counter)
{
    constexpr int S = 32 / LPT;                             // token streams per warp
    constexpr int T = 32 * W;                               // threads per document
    constexpr int R = (4 * LPT * CPL + T - 1) / T;          // K-phase topics per thread
    constexpr int RV = (R % 2 == 0) ? 2 : 1;                // ... owned as RV consecutive topics (LDS.64 / FADD2 owner sums)
#define TOPIC(r) (RV * (tid + T * ((r) / RV)) + ((r) % RV))
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kl = lane % LPT, ts = lane / LPT;
    const int K = p.K, K_ld = p.K_ld, CH = K_ld >> 2, RS = p.RS;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    (void)cap2;

    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw);
    int *next_s = reinterpret_cast<int *>(smem_raw + 16);
    float *csum_s = reinterpret_cast<float *>(smem_raw + 32);        // [W <= 8]
    float *tsum_s = reinterpret_cast<float *>(smem_raw + 64);        // [W <= 8]
    unsigned *dsum_s = reinterpret_cast<unsigned *>(smem_raw + 96);  // [2][W <= 8]
    float *gs = reinterpret_cast<float *>(smem_raw + 256);           // [W][S][RS]
    float *e_s = gs + (size_t)W * S * RS;                            // [RS]
    float *tile = e_s + RS;                                          // [cap][RS]
    float *cnt_s = tile + (size_t)cap * RS;                          // [cap]
    int *term_s = reinterpret_cast<int *>(cnt_s + cap);              // [cap]

    // K-phase state: topics i = TOPIC(r)
    float alpha_k[R], Eold_k[R], Enew_k[R], e_k[R], gam_k[R];
    double esum_k[R];
    float asum = 0.0f;
    for (int i = lane; i < K; i += 32) asum += p.alpha[i];
    asum = warp_sum(asum);
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = TOPIC(r);
        alpha_k[r] = (i < K) ? p.alpha[i] : 0.0f;
        esum_k[r] = 0.0;
        Enew_k[r] = gam_k[r] = Eold_k[r] = e_k[r] = 0.0f;
    }
    // convergence test in fixed point: sum_i (dE_i)^2 * (2^20 / vtol^2) < 2^20, summed with one REDUX per warp
    const float dscale = (p.vtol > 0.0f) ? 1048576.0f / (p.vtol * p.vtol) : 0.0f;
    double elbo_thr = 0.0;
    unsigned long long sweeps_thr = 0;
    unsigned phase = 0;
    if (p.stage_bulk && tid == 0) mbar_init(mbar, 1);
    cta_sync<W>();

    // documents per draw: kDocChunk, fewer when the launch has less than ~4 draws per CTA (multi-GPU shards: balance over atomics)
    const int chunk = max(1, min(kDocChunk, (doc_end - doc_begin) / (4 * (int)gridDim.x)));
    int d_next = 0, d_lim = 0;
    for (;;) {
        // documents are drawn from the bucket's work counter kDocChunk at a time: one same-address atomic per document
        // serialises in L2 (128 804 of them cost ~0.4 ms of a 2.4 ms E-step)
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
        const bool ovf = Nd > cap;

        // stage the document: term ids + counts, then its K x N_d slab of beta -- one TMA bulk copy per term row
        // (a row is K_ld*4 contiguous bytes in HBM/L2), completion tracked by an mbarrier
        float csum = 0.0f;
        for (int n = tid; n < Nd; n += T) {
            const float c = p.counts[o + n];
            csum += c;
            if (n < ns) {
                term_s[n] = p.terms[o + n];
                cnt_s[n] = c;
            }
        }
        csum = warp_sum(csum);
        if (W > 1 && lane == 0) csum_s[warp] = csum;
        if (p.stage_bulk) {
            fence_proxy_async_smem();
            cta_sync<W>();
            if (tid == 0) mbar_arrive_expect_tx(mbar, (unsigned)(ns * K_ld * 4));
            cta_sync<W>();
            for (int n = tid; n < ns; n += T) bulk_g2s(tile + n * RS, p.beta + (size_t)term_s[n] * K_ld, (unsigned)(K_ld * 4), mbar);
        } else {
            cta_sync<W>();
            for (int c = tid; c < ns * CH; c += T) {
                const int n = c / CH, q = c - n * CH;
                cp_async16(tile + n * RS + 4 * q, p.beta + (size_t)term_s[n] * K_ld + 4 * q);
            }
            cp_async_commit();
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = TOPIC(r);
            Eold_k[r] = (i < K) ? p.Elogtheta[(size_t)d * K_ld + i] : 0.0f;
            e_k[r] = (i < K) ? expf(Eold_k[r]) : 0.0f;
            if (i < K_ld) e_s[i] = e_k[r];
        }
        if (W > 1) {
            csum = 0.0f;
#pragma unroll
            for (int w = 0; w < W; w++) csum += csum_s[w];
        }
        // sum(gamma_d) = sum(alpha) + sum_n c_n + K*EPS whatever phi is (each phi column sums to one), so
        // digamma(sum gamma) (LDA.jl:138) is a per-document constant
        const float gsum = (asum + csum) + (float)K * TMVB_EPS;
        const float psi_sum = psi_lgamma<false>(gsum).psi;
        if (p.stage_bulk) {
            mbar_wait(mbar, phase);
            phase ^= 1u;
        } else {
            cp_async_wait_all();
        }
        cta_sync<W>();

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
        ta.r0 = warp;
        ta.rstep = W;

        float4 e[CPL];
        int v = 0;
        for (;;) {
            // ---- token phase: update_phi! + the phi*counts product of update_gamma! (LDA.jl:143-154)
#pragma unroll
            for (int m = 0; m < CPL; m++)
                e[m] = (m < CPL - 1 || kl + LPT * m < CH) ? reinterpret_cast<const float4 *>(e_s)[kl + LPT * m] : zero4;
            float4 g[CPL];
            float tsum = 0.0f;
#pragma unroll
            for (int m = 0; m < CPL; m++) g[m] = zero4;
            if (!ovf)
                tok_sweep<LPT, CPL, false, true, (W == 2 ? 1 : kSweepUnroll)>(ta, ts, kl, e, g, tsum);
            else
                tok_sweep<LPT, CPL, true, true, (W == 2 ? 1 : kSweepUnroll)>(ta, ts, kl, e, g, tsum);
#pragma unroll
            for (int m = 0; m < CPL; m++)
                if (m < CPL - 1 || kl + LPT * m < CH) reinterpret_cast<float4 *>(gs + (warp * S + ts) * RS)[kl + LPT * m] = g[m];
            float tt = across_streams_sum<LPT>(tsum);
            if (W > 1 && lane == 0) tsum_s[warp] = tt;
            cta_sync<W>();  // also orders this sweep's reads of e_s before the K phase overwrites it
            if (W > 1) {
                tt = 0.0f;
#pragma unroll
                for (int w = 0; w < W; w++) tt += tsum_s[w];
            }

            // ---- K phase: update_gamma! (LDA.jl:143-146), update_Elogtheta! (LDA.jl:136-139)
            float dpart = 0.0f;
            float e_new[R];
            float g_own[R];
            if (RV == 2) {
#pragma unroll
                for (int r = 0; r < R; r += 2) {
                    const float2 gg = (TOPIC(r) < K_ld) ? owner_sum2<W * S>(gs, RS, TOPIC(r)) : make_float2(0.f, 0.f);
                    g_own[r] = gg.x;
                    g_own[r + (R > 1 ? 1 : 0)] = gg.y;
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) g_own[r] = (TOPIC(r) < K) ? owner_sum<W * S>(gs, RS, TOPIC(r)) : 0.0f;
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = TOPIC(r);
                const float gi = (i < K) ? g_own[r] : 0.0f;
                // gamma = EPS + (alpha + phi*counts),   phi*counts = e .* g + eps * sum_n t_n
                gam_k[r] = (i < K) ? (alpha_k[r] + fmaf(e_k[r], gi, TMVB_EPS * tt)) + TMVB_EPS : 1.0f;
                Enew_k[r] = psi_lgamma<false, true>(gam_k[r]).psi - psi_sum;
                e_new[r] = 0.0f;
                if (i < K) {
                    const float df = Enew_k[r] - Eold_k[r];
                    dpart = fmaf(df, df, dpart);
                    e_new[r] = fast_exp(Enew_k[r]);
                }
            }
            v++;
            // LDA.jl:175: stop when ||Elogtheta - Elogtheta_old||_2 < vtol (or after viter sweeps)
            if (v >= p.viter) break;
            // per-lane clamp 2^25 keeps the integer sum over up to 64 threads below 2^31
            unsigned dtot = (dscale > 0.0f) ? __reduce_add_sync(0xffffffffu, (unsigned)fminf(dpart * dscale, 33554432.0f)) : 0xffffffffu;
            if (W > 1) {
                // tentatively publish exp(Elogtheta_new): the token phase keeps e in registers, so overwriting e_s is
                // harmless even if the document turns out to have converged
#pragma unroll
                for (int r = 0; r < R; r++)
                    if (TOPIC(r) < K_ld) e_s[TOPIC(r)] = e_new[r];
                if (lane == 0) dsum_s[(v & 1) * W + warp] = dtot;
                __syncthreads();
                dtot = 0;
#pragma unroll
                for (int w = 0; w < W; w++) {
                    const unsigned x = dsum_s[(v & 1) * W + w];
                    dtot = (x > 0x7fffffffu - dtot) ? 0x7fffffffu : dtot + x;
                }
                if (dscale > 0.0f && dtot < 1048576u) break;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    Eold_k[r] = Enew_k[r];
                    e_k[r] = e_new[r];
                }
            } else {
                if (dscale > 0.0f && dtot < 1048576u) break;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    Eold_k[r] = Enew_k[r];
                    e_k[r] = e_new[r];
                    if (TOPIC(r) < K_ld) e_s[TOPIC(r)] = e_new[r];
                }
                __syncwarp();
            }
        }

        // update_beta!(model, d) (LDA.jl:129-132): scatter the last phi, weighted by counts
        if (!(p.dbg & 2)) {
            float ent = 0.0f;
            if (!ovf)
                tok_final<LPT, CPL, false, true, (ELBO ? 2 : 0)>(ta, ts, kl, e, ent);
            else
                tok_final<LPT, CPL, true, true, (ELBO ? 2 : 0)>(ta, ts, kl, e, ent);
            if (ELBO) elbo_thr += (double)ent;
        }

        float a = 0.0f;
        const size_t hrow = p.host_E ? (size_t)__ldg(p.perm + d) * (size_t)K : 0;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = TOPIC(r);
            if (i < K_ld) {
                const bool ok = i < K;
                p.gamma[(size_t)d * K_ld + i] = ok ? gam_k[r] : 0.0f;
                p.Elogtheta[(size_t)d * K_ld + i] = ok ? Enew_k[r] : 0.0f;
                p.Elogtheta_old[(size_t)d * K_ld + i] = ok ? Eold_k[r] : 0.0f;
                if (p.host_E && ok) {
                    __stcs(p.host_gamma + hrow + i, gam_k[r]);
                    __stcs(p.host_E + hrow + i, Enew_k[r]);
                }
                if (ok) {
                    esum_k[r] += (double)Enew_k[r];
                    // lnG(gamma_i), and the Elogtheta_old part of the entropy of the last phi:
                    // sum_n c_n phi_ni ln e_i = (gamma_i - alpha_i) Elogtheta_old_i
                    if (ELBO) a += psi_lgamma<true>(gam_k[r]).lg - (gam_k[r] - alpha_k[r]) * Eold_k[r];
                }
            }
        }
        // Dirichlet entropy (utils.jl:163-180) + Elogpz (LDA.jl:57-60): with gamma = alpha + phi*c and
        // psi(gamma_i) = Elogtheta_i + psi(sum gamma) they collapse to
        //   sum_i lnG(gamma_i) - lnG(sum gamma) + sum_i (1 - alpha_i) Elogtheta_i ;
        // the last sum is linear in sum_d Elogtheta_d and is added by lda_alpha_kernel in fp64.
        // -Elogqz (LDA.jl:76-79) = sum_n c_n H(phi_n) with phi_ni = u_ni / s_n, ln u_ni = ln beta_old_i,w + Elogtheta_old_i:
        //   sum_n c_n ln s_n  [tok_final]  - sum_i (gamma_i - alpha_i) Elogtheta_old_i  [above]
        //   - sum_ij S_ij ln beta_old_ij  [normalize_kernel, over the reduced statistics]
        if (ELBO) {
            if (tid == 0) a -= psi_lgamma<true>(gsum).lg;
            elbo_thr += (double)a;
        }
        if (tid == 0) sweeps_thr += (unsigned long long)v;
        cta_sync<W>();  // the tile, e_s and gs are free for the next document
    }

    // flush the accumulators
    if (ELBO) {
        const double tot = warp_sum_d(elbo_thr);
        if (lane == 0 && tot != 0.0) atomicAdd(p.small + K_ld, tot);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = TOPIC(r);
        if (i < K && esum_k[r] != 0.0) atomicAdd(p.small + i, esum_k[r]);
    }
    if (tid == 0 && sweeps_thr) atomicAdd(p.small + K_ld + 1, (double)sweeps_thr);
}
#undef TOPIC

// ------------------------------------------------------------------ hybrid E-step (registers + tile) ----------
// lda_estep_hyb_kernel<LPT, CPL, KLD, W, NR, TILE, ELBO>: W warps share a document.  Warp w keeps its NR register rounds
// (tokens ((j W + w) S + ts), j < NR -- RegDoc) for all sweeps; with TILE, tokens beyond W * NR * S live in a shared-memory
// tile (tile round q belongs to warp q % W), staged once per document with one TMA bulk copy per term row.  Design points,
// each from the per-instruction profile of its predecessor (profiles/r2_lda_*):
//   * K_ld is a template parameter: every shared-memory exchange addresses with immediate offsets;
//   * exp(Elogtheta) is dead during the K phase (the scatter pass reloads it from a double-buffered shared copy), so the
//     owner sums have the registers to issue their 16 LDS before the first FADD2 (they were serialised on one register pair);
//   * K_ld <= 64, W <= 2 ("RK"): EVERY warp runs the K phase for all topics, redundantly, so a sweep has ONE CTA barrier (the
//     exchange of the per-warp sums) and no warp idles while another evaluates digamma; a document starts and ends without
//     a barrier (sum_n c_n comes precomputed, the exchange buffers rotate with the document parity);
//   * otherwise (W = 4, or K_ld > 64) the K phase is spread over all threads of the CTA (thread t owns topics 2t, 2t+1): this
//     is what lets K = 200 leave the pure tile kernel -- W = 4, LPT = 8, 48 tokens in registers, a tile less than half the size;
//   * sum_n t_n (the epsilon term of gamma) rides through the exchange as one more column instead of a shuffle reduction;
//   * the K phase uses bare MUFU forms; the scatter pass reuses the normalisers s_n of the last sweep for the register rounds.
// Shared memory: header (128 B) | gs [W][S][RSG] | xs [RK ? 4 : 2][W][RSG] | e_s [RK ? W : 1][2][RS] | tile [cap][RS] | cnt_s [cap] |
// term_s [cap]   (RSG = row stride with room for one more column: the per-warp sum_n t_n travels in it).
__host__ __device__ constexpr int hyb_row_stride(int CH, int lpt)   // row_stride() of tmvb_shard.cu as a constant expression
{
    int r = CH;
    if (lpt < 8)
        while (r % (2 * lpt) != lpt) r++;
    return 4 * r;
}
static size_t lda_hyb_fixed_smem(int K_ld, int lpt, int W)
{
    const size_t RS = hyb_row_stride(K_ld / 4, lpt), RSG = hyb_row_stride(K_ld / 4 + 1, lpt);
    const bool rk = W <= 2 && K_ld <= 62;   // RK of the kernel: exchange buffers x 4 (document parity), one e copy per warp
    return 128 + (size_t)W * (32 / lpt) * RSG * 4 + (size_t)(rk ? 4 : 2) * W * RSG * 4 + (size_t)2 * (rk ? W : 1) * RS * 4;
}

// registers per thread the hybrid variants are held to: 65536 / (32 * resident warps per SM) in units of 8 -- 12 / 10 / 8
constexpr int lda_hyb_maxreg(int NR) { return NR <= 3 ? 168 : NR == 4 ? 200 : 255; }

// owner-lane sums of topics i, i+1 over the S per-stream partials: all loads first, then a tree of FADD2
template <int S>
__device__ __forceinline__ float2 owner_sum2_tree(const float *gs, int RS, int i)
{
    f32x2 v[S];
#pragma unroll
    for (int w = 0; w < S; w++) v[w] = *reinterpret_cast<const f32x2 *>(gs + w * RS + i);
#pragma unroll
    for (int h = S / 2; h > 0; h >>= 1)
#pragma unroll
        for (int w = 0; w < h; w++) v[w] = add2(v[w], v[w + h]);
    float2 r;
    unpk2(v[0], r.x, r.y);
    return r;
}
// developer switches for A/B builds of the hybrid kernel (tools/build_variants.sh); the defaults are the shipped configuration
#ifndef TMVB_V_TSUMCOL
#define TMVB_V_TSUMCOL 0   // 1: sum_n t_n rides through the exchange as an extra column; 0: shuffle reduction in the token phase
#endif
#ifndef TMVB_V_STS64
#define TMVB_V_STS64 0     // 1: per-stream partials stored as 8-byte pairs; 0: as 16-byte chunks
#endif
#ifndef TMVB_V_OPAQUE
#define TMVB_V_OPAQUE 0    // 1: thread index and shared-window base are read once and kept (no S2R / S2UR rematerialisation in the sweep loop)
#endif
#ifndef TMVB_V_NODOCSYNC
#define TMVB_V_NODOCSYNC 1 // 1: RK variants without a tile start and end a document without a CTA barrier
#endif
__device__ __forceinline__ void sts64(unsigned addr, f32x2 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }

template <int LPT, int CPL, int KLD, int W, int NR, bool TILE, bool ELBO>
__global__ void __maxnreg__(lda_hyb_maxreg(NR)) lda_estep_hyb_kernel(const LdaDev p, int doc_begin, int doc_end, int cap, int cap2, int *counter)
{
    constexpr int S = 32 / LPT;
    constexpr int T = 32 * W;
    constexpr int K_ld = KLD, CH = KLD / 4, RS = hyb_row_stride(KLD / 4, LPT), RSG = hyb_row_stride(KLD / 4 + 1, LPT);
    constexpr int PM = (KLD + 2 + 63) / 64;       // topic pairs (+ the sum_n t_n column) per lane when a warp folds its own S streams
    constexpr int REGTOK = W * NR * S;            // tokens of a document that live in registers
    constexpr bool RK = (W <= 2 && KLD <= 62);    // every warp runs the whole K phase (one barrier per sweep, none per document)
    constexpr bool DOCSYNC = TILE || !RK || !TMVB_V_NODOCSYNC;   // a document starts and ends with a CTA barrier
    static_assert(KLD % 8 == 0 && KLD <= 4 * LPT * CPL && KLD > 4 * LPT * (CPL - 1), "K_ld must fill the lane layout's last chunk column");
    static_assert(64 * W >= KLD + 2, "the K phase keeps one topic pair (or the sum_n t_n column) per thread");
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    int tid_ = threadIdx.x;
    unsigned char *smem_raw = smem_dyn;
#if TMVB_V_OPAQUE
    // opaque copies: the compiler otherwise re-reads SR_TID and rebuilds the shared-window base (S2UR SR_CgaCtaId, ~30 cycles of
    // exposed latency each) several times per sweep instead of keeping them in registers
    asm volatile("mov.u32 %0, %0;" : "+r"(tid_));
    {
        unsigned long long g = reinterpret_cast<unsigned long long>(smem_dyn);
        asm volatile("mov.u64 %0, %0;" : "+l"(g));
        smem_raw = reinterpret_cast<unsigned char *>(g);
    }
#endif
    const int tid = tid_, lane = tid & 31, warp = tid >> 5;
    const int kl = lane % LPT, ts = lane / LPT;
    const int K = p.K;
    (void)cap2;
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw);
    int *next_s = reinterpret_cast<int *>(smem_raw + 16);
    float *asum_s = reinterpret_cast<float *>(smem_raw + 32);        // [W <= 4]
    unsigned *dsum_s = reinterpret_cast<unsigned *>(smem_raw + 64);  // [2][W <= 4]
    float *gs = reinterpret_cast<float *>(smem_raw + 128);           // [W][S][RSG]
    float *xs = gs + (size_t)W * S * RSG;                            // [RK ? 4 : 2][W][RSG]
    float *e_s = xs + (size_t)(RK ? 4 : 2) * W * RSG;                // [RK ? W : 1][2][RS]
    float *tile = e_s + (size_t)2 * (RK ? W : 1) * RS;               // [cap][RS]
    float *cnt_s = tile + (size_t)cap * RS;                          // [cap]
    int *term_s = reinterpret_cast<int *>(cnt_s + cap);              // [cap]
    float *gs_w = gs + (size_t)warp * S * RSG;
    float *e_w = e_s + (size_t)(RK ? warp : 0) * 2 * RS;
    // shared-window address of this lane's slice of its stream's partial vector (8-byte stores: an STS.128 would cost four MOVs
    // per chunk to line the two accumulator pairs up in one aligned register quad)
    const unsigned gs_st = (unsigned)__cvta_generic_to_shared(gs_w + ts * RSG + 4 * kl);

    // K-phase ownership: topics i0, i0 + 1 -- pair `lane` in every warp (RK), else pair `tid`; pair K_ld / 2 is the sum_n t_n column
    const int i0 = 2 * (RK ? lane : tid);
    const bool in_ld = i0 < K_ld, in_x = i0 < K_ld + 2, ok0 = i0 < K, ok1 = i0 + 1 < K;
    const bool writer = RK ? (warp == 0) : true;                     // who stores the document's K-vectors and accumulates its sums
    const float a0 = ok0 ? p.alpha[i0] : 0.0f, a1 = ok1 ? p.alpha[i0 + 1] : 0.0f;
    float asum = warp_sum(a0 + a1);
    if (!RK) {
        if (lane == 0) asum_s[warp] = asum;
        __syncthreads();
        asum = 0.0f;
#pragma unroll
        for (int w = 0; w < W; w++) asum += asum_s[w];
    }
    // convergence test in fixed point: sum_i dE_i^2 * (2^20 / vtol^2) < 2^20 with one REDUX per warp.  vtol = 0: the scale is
    // +inf, every term saturates at the clamp (0 * inf = NaN is dropped by fminf) and the test never passes, as ||.|| < 0 never does
    const float dscale = (p.vtol > 0.0f) ? 1048576.0f / (p.vtol * p.vtol) : __int_as_float(0x7f800000);
    const float m0 = ok0 ? 1.0f : 0.0f, m1 = ok1 ? 1.0f : 0.0f;     // pad topics do not count in the test
    const f32x2 keps_init = pk2(kl == 0 ? (float)K * TMVB_EPS : 0.0f, 0.0f);
    const ulonglong2 zero = make_ulonglong2(0ull, 0ull);
    double esum0 = 0.0, esum1 = 0.0, elbo_thr = 0.0;
    unsigned long long sweeps_thr = 0;
    unsigned phase = 0;
    if (TILE && tid == 0) mbar_init(mbar, 1);
    cta_sync<W>();

    const int chunk = max(1, min(kDocChunk, (doc_end - doc_begin) / (4 * (int)gridDim.x)));
    int d_next = 0, d_lim = 0, dpar = 0;
    long long o_cur = 0, o_end = 0;
    for (;;) {
        if (d_next >= d_lim) {  // kDocChunk documents per draw from the work counter (see lda_estep_kernel)
            if (W == 1) {
                if (lane == 0) d_next = doc_begin + atomicAdd(counter, chunk);
                d_next = __shfl_sync(0xffffffffu, d_next, 0);
            } else {
                if (tid == 0) *next_s = doc_begin + atomicAdd(counter, chunk);
                __syncthreads();
                d_next = *next_s;
            }
            d_lim = min(d_next + chunk, doc_end);
            if (d_next < doc_end) {
                o_cur = p.doc_off[d_next];
                o_end = p.doc_off[d_next + 1];
            }
        }
        if (d_next >= doc_end) break;
        const int d = d_next++;
        const long long o = o_cur;
        const int Nd = (int)(o_end - o);
        o_cur = o_end;
        if (d_next < d_lim) o_end = p.doc_off[d_next + 1];
        const int n_tile = TILE ? max(Nd - REGTOK, 0) : 0;   // <= cap: the launch buckets are planned that way
        const float csum = __ldg(p.doc_c + d);

        RegDoc<LPT, CPL, NR> rd;
        reg_load<LPT, CPL, W, NR>(rd, p.beta, p.terms + o, p.counts + o, Nd, K_ld, warp, ts, kl, p.dbg);
        // the documents of a draw are adjacent in the CSR arrays and in Elogtheta: pull the next one's lines into L2 now
        // (its term ids come from HBM; the dependent row gather cannot start before they arrive)
        if (d_next < d_lim && warp == 0 && lane < 12 && !(p.dbg & 8)) {
            const char *a = lane < 4 ? reinterpret_cast<const char *>(p.terms + o_cur) + 128 * lane
                            : lane < 8 ? reinterpret_cast<const char *>(p.counts + o_cur) + 128 * (lane - 4)
                                       : reinterpret_cast<const char *>(p.Elogtheta + (size_t)(d + 1) * K_ld) + 128 * (lane - 8);
            if (lane < 8 || 128 * (lane - 8) < K_ld * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
        // tile: ids and counts of the tokens beyond the register rounds, then one TMA bulk copy per term row
        if (n_tile > 0) {
            for (int n = tid; n < n_tile; n += T) {
                term_s[n] = __ldg(p.terms + o + REGTOK + n);
                cnt_s[n] = __ldg(p.counts + o + REGTOK + n);
            }
            fence_proxy_async_smem();
            cta_sync<W>();
            if (tid == 0) mbar_arrive_expect_tx(mbar, (unsigned)(n_tile * K_ld * 4));
            cta_sync<W>();
            for (int n = tid; n < n_tile; n += T) bulk_g2s(tile + n * RS, p.beta + (size_t)term_s[n] * K_ld, (unsigned)(K_ld * 4), mbar);
        }

        float Eo0 = 0.0f, Eo1 = 0.0f, En0 = 0.0f, En1 = 0.0f, ek0 = 0.0f, ek1 = 0.0f, gam0 = 1.0f, gam1 = 1.0f;
        if (in_ld) {
            const float2 E = *reinterpret_cast<const float2 *>(p.Elogtheta + (size_t)d * K_ld + i0);
            Eo0 = ok0 ? E.x : 0.0f;
            Eo1 = ok1 ? E.y : 0.0f;
            ek0 = ok0 ? expf(Eo0) : 0.0f;
            ek1 = ok1 ? expf(Eo1) : 0.0f;
            *reinterpret_cast<float2 *>(e_w + i0) = make_float2(ek0, ek1);
        }
        if (n_tile > 0) {
            mbar_wait(mbar, phase);
            phase ^= 1u;
        }
        if (DOCSYNC)
            cta_sync<W>();
        else
            __syncwarp();
        // sum(gamma_d) = sum(alpha) + sum_n c_n + K*EPS whatever phi is: digamma(sum gamma) (LDA.jl:138) is a per-document constant
        const float gsum = (asum + csum) + (float)K * TMVB_EPS;
        const float psi_sum = psi_lgamma<false>(gsum).psi;
        // this document's pair of exchange buffers: they alternate with the documents THIS CTA processes, so a warp that runs
        // ahead into the next document never writes a buffer the other warp may still be reading
        dpar ^= 1;
        float *xd = xs + (size_t)(RK ? dpar : 0) * 2 * W * RSG;

        float s_keep[NR];
        int v = 0;
        for (;;) {
            const int eb = v & 1;
            float tt_w = 0.0f;   // this warp's sum_n t_n
            {
                // ---- token phase: update_phi! + the phi*counts product of update_gamma! (LDA.jl:143-154)
                f32x2 e01[CPL], e23[CPL];
#pragma unroll
                for (int m = 0; m < CPL; m++) {
                    const ulonglong2 ev = (m < CPL - 1 || kl + LPT * m < CH) ? reinterpret_cast<const ulonglong2 *>(e_w + eb * RS)[kl + LPT * m] : zero;
                    e01[m] = ev.x;
                    e23[m] = ev.y;
                }
                f32x2 g01[CPL], g23[CPL];
                float tsum = 0.0f;
#pragma unroll
                for (int m = 0; m < CPL; m++) g01[m] = g23[m] = 0ull;
                if (n_tile > 0) tile_sweep_pk<LPT, CPL, true>(tile, cnt_s, n_tile, RS, keps_init, CH, ts, kl, warp, W, e01, e23, g01, g23, tsum);
                reg_sweep_keep<LPT, CPL, NR, true>(rd, keps_init, e01, e23, g01, g23, tsum, s_keep);
#pragma unroll
                for (int m = 0; m < CPL; m++)
                    if (m < CPL - 1 || kl + LPT * m < CH) {
#if TMVB_V_STS64
                        sts64(gs_st + 16 * LPT * m, g01[m]);
                        sts64(gs_st + 16 * LPT * m + 8, g23[m]);
#else
                        reinterpret_cast<ulonglong2 *>(gs_w + ts * RSG)[kl + LPT * m] = make_ulonglong2(g01[m], g23[m]);
#endif
                    }
#if TMVB_V_TSUMCOL
                if (kl == 0) gs_w[ts * RSG + K_ld] = tsum;   // the extra column: this stream's sum_n t_n
#else
                tt_w = across_streams_sum<LPT>(tsum);
#endif
            }
            __syncwarp();
            // ---- fold the S per-stream partials of this warp, then the W warps, into the K-phase layout
            float2 gg = make_float2(0.f, 0.f);
            if (W == 1) {
                if (TMVB_V_TSUMCOL ? in_x : in_ld) gg = owner_sum2_tree<S>(gs_w, RSG, i0);
            } else {
                float *xb = xd + (size_t)eb * W * RSG;
#pragma unroll
                for (int m = 0; m < PM; m++) {
                    const int i = 2 * (lane + 32 * m);
#if TMVB_V_TSUMCOL
                    if (i < K_ld + 2) *reinterpret_cast<float2 *>(xb + warp * RSG + i) = owner_sum2_tree<S>(gs_w, RSG, i);
#else
                    if (i < K_ld) *reinterpret_cast<float2 *>(xb + warp * RSG + i) = owner_sum2_tree<S>(gs_w, RSG, i);
#endif
                }
#if !TMVB_V_TSUMCOL
                if (lane == 0) xb[warp * RSG + K_ld] = tt_w;
#endif
                __syncthreads();
                if (TMVB_V_TSUMCOL ? in_x : in_ld) {
#pragma unroll
                    for (int w = 0; w < W; w++) {
                        const float2 x = *reinterpret_cast<const float2 *>(xb + w * RSG + i0);
                        gg.x += x.x;
                        gg.y += x.y;
                    }
                }
            }
            float tt;
#if TMVB_V_TSUMCOL
            // sum_n t_n lives in the first slot of pair K_ld / 2
            if (RK || W == 1) {
                tt = __shfl_sync(0xffffffffu, gg.x, K_ld / 2);
            } else {
                tt = 0.0f;
                const float *xb = xd + (size_t)eb * W * RSG;
#pragma unroll
                for (int w = 0; w < W; w++) tt += xb[w * RSG + K_ld];
            }
#else
            if (W == 1) {
                tt = tt_w;
            } else {
                tt = 0.0f;
                const float *xb = xd + (size_t)eb * W * RSG;
#pragma unroll
                for (int w = 0; w < W; w++) tt += xb[w * RSG + K_ld];   // written below, before the barrier
            }
#endif
            // ---- K phase: update_gamma! (LDA.jl:143-146), update_Elogtheta! (LDA.jl:136-139)
            v++;
            // gamma = EPS + (alpha + phi*counts),   phi*counts = e .* g + eps * sum_n t_n   (pad topics: gamma = 1, harmless)
            const float et = TMVB_EPS * tt;
            gam0 = ok0 ? (a0 + fmaf(ek0, gg.x, et)) + TMVB_EPS : 1.0f;
            gam1 = ok1 ? (a1 + fmaf(ek1, gg.y, et)) + TMVB_EPS : 1.0f;
            psi_pair_fast(gam0, gam1, En0, En1);
            En0 -= psi_sum;
            En1 -= psi_sum;
            if (v >= p.viter) break;
            // exp(Elogtheta_new); pad topics need no mask: their table entries are zero, so they reach neither s_n nor g
            const float en0 = ex2_ftz(En0 * 1.4426950408889634f), en1 = ex2_ftz(En1 * 1.4426950408889634f);
            // LDA.jl:175: stop when ||Elogtheta - Elogtheta_old||_2 < vtol; fixed point, one REDUX per warp
            const float d0 = (En0 - Eo0) * m0, d1 = (En1 - Eo1) * m1;
            const float dpart = fmaf(d0, d0, d1 * d1);
            // per-thread clamp 2^23 keeps the integer sum over up to 128 threads below 2^31
            unsigned dtot = __reduce_add_sync(0xffffffffu, (unsigned)fminf(dpart * dscale, 8388608.0f));
            if (RK || W == 1) {
                // every warp holds all topic pairs: the decision is warp-local and identical in all warps
                if (dtot < 1048576u) break;
                if (in_ld) *reinterpret_cast<float2 *>(e_w + (eb ^ 1) * RS + i0) = make_float2(en0, en1);
                __syncwarp();
            } else {
                // publish exp(Elogtheta_new) into the other buffer (the scatter pass reloads this sweep's e from buffer eb)
                if (in_ld) *reinterpret_cast<float2 *>(e_w + (eb ^ 1) * RS + i0) = make_float2(en0, en1);
                if (lane == 0) dsum_s[eb * W + warp] = dtot;
                __syncthreads();
                dtot = 0;
#pragma unroll
                for (int w = 0; w < W; w++) {
                    const unsigned x = dsum_s[eb * W + w];
                    dtot = (x > 0x7fffffffu - dtot) ? 0x7fffffffu : dtot + x;
                }
                if (dtot < 1048576u) break;
            }
            Eo0 = En0;
            Eo1 = En1;
            ek0 = en0;
            ek1 = en1;
        }

        // update_beta!(model, d) (LDA.jl:129-132): scatter the last phi, weighted by counts (e of the last sweep: buffer (v-1) & 1)
        if (!(p.dbg & 2)) {
            const int eb = (v - 1) & 1;
            f32x2 e01[CPL], e23[CPL];
#pragma unroll
            for (int m = 0; m < CPL; m++) {
                const ulonglong2 ev = (m < CPL - 1 || kl + LPT * m < CH) ? reinterpret_cast<const ulonglong2 *>(e_w + eb * RS)[kl + LPT * m] : zero;
                e01[m] = ev.x;
                e23[m] = ev.y;
            }
            float ent = 0.0f;
            reg_final_keep<LPT, CPL, NR, true, ELBO>(rd, p.stats, K, K_ld, kl, e01, e23, s_keep, ent, p.dbg);
            if (n_tile > 0) tile_final_pk<LPT, CPL, true, ELBO>(tile, cnt_s, term_s, p.stats, n_tile, RS, K, K_ld, CH, ts, kl, warp, W, e01, e23, ent, p.dbg);
            if (ELBO) elbo_thr += (double)ent;
        }
        if (writer) {
            if (in_ld) {
                *reinterpret_cast<float2 *>(p.gamma + (size_t)d * K_ld + i0) = make_float2(ok0 ? gam0 : 0.0f, ok1 ? gam1 : 0.0f);
                *reinterpret_cast<float2 *>(p.Elogtheta + (size_t)d * K_ld + i0) = make_float2(ok0 ? En0 : 0.0f, ok1 ? En1 : 0.0f);
                *reinterpret_cast<float2 *>(p.Elogtheta_old + (size_t)d * K_ld + i0) = make_float2(ok0 ? Eo0 : 0.0f, ok1 ? Eo1 : 0.0f);
            }
            if (p.host_E) {   // uniform: the row goes over the bus while the other documents are swept
                const size_t hrow = (size_t)__ldg(p.perm + d) * (size_t)K + i0;
                if (ok1 && !(K & 1)) {   // rows of an even K start 8-byte aligned
                    __stcs(reinterpret_cast<float2 *>(p.host_gamma + hrow), make_float2(gam0, gam1));
                    __stcs(reinterpret_cast<float2 *>(p.host_E + hrow), make_float2(En0, En1));
                } else {
                    if (ok0) {
                        __stcs(p.host_gamma + hrow, gam0);
                        __stcs(p.host_E + hrow, En0);
                    }
                    if (ok1) {
                        __stcs(p.host_gamma + hrow + 1, gam1);
                        __stcs(p.host_E + hrow + 1, En1);
                    }
                }
            }
            float a = 0.0f;
            if (ok0) {
                esum0 += (double)En0;
                if (ELBO) a += psi_lgamma<true>(gam0).lg - (gam0 - a0) * Eo0;
            }
            if (ok1) {
                esum1 += (double)En1;
                if (ELBO) a += psi_lgamma<true>(gam1).lg - (gam1 - a1) * Eo1;
            }
            // the per-document ELBO terms: see lda_estep_kernel
            if (ELBO) {
                if (tid == 0) a -= psi_lgamma<true>(gsum).lg;
                elbo_thr += (double)a;
            }
            if (tid == 0) sweeps_thr += (unsigned long long)v;
        }
        // the tile, the shared e_s and the header slots are free for the next document; with RK and no tile every buffer a warp
        // writes before the next exchange barrier is its own (gs_w, e_w) or belongs to the other document parity (xd)
        if (DOCSYNC)
            cta_sync<W>();
        else
            __syncwarp();
    }

    if (ELBO) {
        const double tot = warp_sum_d(elbo_thr);
        if (lane == 0 && tot != 0.0) atomicAdd(p.small + K_ld, tot);
    }
    if (writer) {
        if (ok0 && esum0 != 0.0) atomicAdd(p.small + i0, esum0);
        if (ok1 && esum1 != 0.0) atomicAdd(p.small + i0 + 1, esum1);
    }
    if (tid == 0 && sweeps_thr) atomicAdd(p.small + K_ld + 1, (double)sweeps_thr);
}

}  // namespace tmvb
