/*
 * oracle/ctpf_oracle.c -- fp64 CPU restatement of the reference's CPU collaborative topic Poisson factorization
 * model (src/CTPF.jl).
 *
 * TEST INFRASTRUCTURE ONLY (see lda_oracle.c).  PARITY UNPINNED: no golden vectors exist in the reference and it
 * cannot run here; cross-checked against oracle/numpy_twin.py (CTPFTwin), which evaluates the ELBO through the
 * closed form the Binomial/log-Gamma sums collapse to, so agreement also checks that identity.
 *
 * Layout: alef[K*j + i] (K x V column-major), he[K*u + i] (K x U), gimel[K*d + i], zayin[K*d + i]; bet, vav,
 * dalet, het of length K.  terms / readers are 0-based.  hyp = {a, b, c, d, e, f, g, h} (CTPF.jl:81).
 *
 * Third-party arithmetic restated (Distributions.jl 0.23, not vendored):
 *   pdf(Binomial(n, p), k) = exp(binomlogpdf) with the package's own override (utils.jl:159-160);
 *   entropy(Multinomial(n, p)) = -lnG(n+1) + n H(p) + sum_i sum_{x=0..n} pdf(Binomial(n, p_i), x) lnG(x+1);
 *   entropy(Gamma(alpha, theta)) = alpha + ln theta + lnG(alpha) + (1 - alpha) psi(alpha).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "special.h"

static double xlogy(double x, double y) { return x != 0.0 ? x * log(y) : 0.0; } /* utils.jl:159 */

static double binompdf(double n, double p, double k) /* utils.jl:160 */
{
    return exp(orc_lgamma(n + 1.0) - orc_lgamma(k + 1.0) - orc_lgamma(n - k + 1.0) + xlogy(k, p) + xlogy(n - k, 1.0 - p));
}

/* sum_{x=0..n} pdf(Binomial(n, p), x) lnG(x+1)  -- the term of CTPF.jl:116,127,138 and of entropy(Multinomial) */
static double binom_lgamma_sum(int64_t n, double p)
{
    double s = 0.0;
    for (int64_t x = 0; x <= n; x++) s += binompdf((double)n, p, (double)x) * orc_lgamma((double)x + 1.0);
    return s;
}

static double entropy_multinomial(int64_t n, int64_t k, const double *p)
{
    double h = 0.0, s;
    for (int64_t i = 0; i < k; i++) if (p[i] > 0.0) h -= p[i] * log(p[i]);
    s = -orc_lgamma((double)n + 1.0) + (double)n * h;
    for (int64_t i = 0; i < k; i++) s += binom_lgamma_sum(n, p[i]);
    return s;
}

static double entropy_gamma(double alpha, double theta) { return alpha + log(theta) + orc_lgamma(alpha) + (1.0 - alpha) * orc_digamma(alpha); }

/* additive_logistic over the rows of a column (utils.jl:114-123): x[0..n) -> softmax in place */
static void softmax(int64_t n, double *x)
{
    double mx = -INFINITY, s = 0.0;
    for (int64_t i = 0; i < n; i++) if (x[i] > mx) mx = x[i];
    for (int64_t i = 0; i < n; i++) { x[i] = exp(x[i] - mx); s += x[i]; }
    for (int64_t i = 0; i < n; i++) x[i] /= s;
}

/* CTPF.jl:327-330 update_phi! */
static void ctpf_phi(int64_t K, int64_t Nd, const int64_t *terms, const double *alef, const double *gimel_d, const double *dalet,
                     const double *bet, double *phi)
{
    for (int64_t n = 0; n < Nd; n++) {
        double *p = phi + K * n;
        const double *a = alef + K * terms[n];
        for (int64_t i = 0; i < K; i++) p[i] = orc_digamma(gimel_d[i]) - log(dalet[i]) - log(bet[i]) + orc_digamma(a[i]);
        softmax(K, p);
    }
}

/* CTPF.jl:334-337 update_xi! */
static void ctpf_xi(int64_t K, int64_t Rd, const int64_t *readers, const double *he, const double *gimel_d, const double *zayin_d,
                    const double *dalet, const double *het, const double *vav, double *xi)
{
    for (int64_t r = 0; r < Rd; r++) {
        double *x = xi + 2 * K * r;
        const double *hh = he + K * readers[r];
        for (int64_t i = 0; i < K; i++) {
            double ph = orc_digamma(hh[i]);
            x[i] = orc_digamma(gimel_d[i]) - log(dalet[i]) - log(vav[i]) + ph;
            x[K + i] = orc_digamma(zayin_d[i]) - log(het[i]) - log(vav[i]) + ph;
        }
        softmax(2 * K, x);
    }
}

typedef struct {
    int64_t K, M, V, U;
    const int64_t *N_cumsum, *terms, *counts, *R_cumsum, *readers, *ratings;
    const double *hyp;
} ctpf_corp;

/* CTPF.jl:232-247 update_elbo!: phi / xi rebuilt from the *_old copies, every expectation with the current values */
double orc_ctpf_elbo(int64_t K, int64_t M, int64_t V, int64_t U, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts,
                     const int64_t *R_cumsum, const int64_t *readers, const int64_t *ratings, const double *hyp, const double *alef,
                     const double *alef_old, const double *he, const double *he_old, const double *bet, const double *bet_old,
                     const double *vav, const double *vav_old, const double *gimel, const double *gimel_old, const double *zayin,
                     const double *zayin_old, const double *dalet, const double *dalet_old, const double *het, const double *het_old,
                     int nthreads)
{
    const double a = hyp[0], b = hyp[1], c = hyp[2], dd = hyp[3], e = hyp[4], f = hyp[5], g = hyp[6], h = hyp[7];
    double *alsum = (double *)calloc(K, sizeof(double)), *hesum = (double *)calloc(K, sizeof(double));
    /* Elogpbeta - Elogqbeta (CTPF.jl:143-150,197-203) */
    double elbo = (double)V * (double)K * (a * log(b) - orc_lgamma(a));
    for (int64_t j = 0; j < V; j++)
        for (int64_t i = 0; i < K; i++) {
            double x = alef[K * j + i];
            alsum[i] += x;
            elbo += (a - 1.0) * (orc_digamma(x) - log(bet[i])) - b * x / bet[i] + entropy_gamma(x, 1.0 / bet[i]);
        }
    /* Elogpeta - Elogqeta (CTPF.jl:161-168,215-221) */
    elbo += (double)U * (double)K * (e * log(f) - orc_lgamma(e));
    for (int64_t u = 0; u < U; u++)
        for (int64_t i = 0; i < K; i++) {
            double x = he[K * u + i];
            hesum[i] += x;
            elbo += (e - 1.0) * (orc_digamma(x) - log(vav[i])) - f * x / vav[i] + entropy_gamma(x, 1.0 / vav[i]);
        }
    int64_t mxn = 1, mxr = 1;
    for (int64_t d = 0; d < M; d++) {
        if (N_cumsum[d + 1] - N_cumsum[d] > mxn) mxn = N_cumsum[d + 1] - N_cumsum[d];
        if (R_cumsum[d + 1] - R_cumsum[d] > mxr) mxr = R_cumsum[d + 1] - R_cumsum[d];
    }
    if (nthreads < 1) nthreads = 1;
    double docs = 0.0;
#pragma omp parallel num_threads(nthreads) reduction(+ : docs)
    {
        double *phi = (double *)malloc(sizeof(double) * K * mxn), *xi = (double *)malloc(sizeof(double) * 2 * K * mxr);
#pragma omp for schedule(dynamic, 16)
        for (int64_t d = 0; d < M; d++) {
            int64_t o = N_cumsum[d], Nd = N_cumsum[d + 1] - o, ro = R_cumsum[d], Rd = R_cumsum[d + 1] - ro;
            const double *gm = gimel + K * d, *zy = zayin + K * d;
            ctpf_phi(K, Nd, terms + o, alef_old, gimel_old + K * d, dalet_old, bet_old, phi);
            ctpf_xi(K, Rd, readers + ro, he_old, gimel_old + K * d, zayin_old + K * d, dalet_old, het_old, vav_old, xi);
            double x = 0.0;
            for (int64_t i = 0; i < K; i++) {
                x -= gm[i] / (dalet[i] * vav[i]) * hesum[i];   /* Elogpya head, CTPF.jl:112 */
                x -= zy[i] / (het[i] * vav[i]) * hesum[i];     /* Elogpyb head, CTPF.jl:123 */
                x -= gm[i] / (dalet[i] * bet[i]) * alsum[i];   /* Elogpz head, CTPF.jl:134 */
            }
            for (int64_t r = 0; r < Rd; r++) {
                const double *xr = xi + 2 * K * r;
                const double *hh = he + K * readers[ro + r];
                int64_t ra = ratings[ro + r];
                for (int64_t i = 0; i < K; i++) {
                    x += (double)ra * xr[i] * (orc_digamma(gm[i]) - log(dalet[i]) + orc_digamma(hh[i]) - log(vav[i])) - binom_lgamma_sum(ra, xr[i]);
                    x += (double)ra * xr[K + i] * (orc_digamma(zy[i]) - log(het[i]) + orc_digamma(hh[i]) - log(vav[i])) - binom_lgamma_sum(ra, xr[K + i]);
                }
                x += entropy_multinomial(ra, 2 * K, xr);       /* -Elogqy, CTPF.jl:179-185 */
            }
            for (int64_t n = 0; n < Nd; n++) {
                const double *p = phi + K * n;
                const double *al = alef + K * terms[o + n];
                int64_t cn = counts[o + n];
                for (int64_t i = 0; i < K; i++)
                    x += (double)cn * p[i] * (orc_digamma(gm[i]) - log(dalet[i]) + orc_digamma(al[i]) - log(bet[i])) - binom_lgamma_sum(cn, p[i]);
                x += entropy_multinomial(cn, K, p);            /* -Elogqz, CTPF.jl:188-194 */
            }
            x += (double)K * (c * log(dd) - orc_lgamma(c)) + (double)K * (g * log(h) - orc_lgamma(g));
            for (int64_t i = 0; i < K; i++) {
                x += (c - 1.0) * (orc_digamma(gm[i]) - log(dalet[i])) - dd * gm[i] / dalet[i];   /* Elogptheta, CTPF.jl:152-159 */
                x += (g - 1.0) * (orc_digamma(zy[i]) - log(het[i])) - h * zy[i] / het[i];        /* Elogpepsilon, CTPF.jl:170-177 */
                x += entropy_gamma(gm[i], 1.0 / dalet[i]) + entropy_gamma(zy[i], 1.0 / het[i]);  /* CTPF.jl:206-212,224-230 */
            }
            docs += x;
        }
        free(phi);
        free(xi);
    }
    free(alsum);
    free(hesum);
    return elbo + docs;
}

/*
 * CTPF.jl:344-402 train! (without the recommendation post-processing :373-400).
 * In/out: alef[K*V], he[K*U], bet, vav, dalet, het [K], gimel[K*M], zayin[K*M].
 * Out: the *_old copies, elbo_trace[iter+1], sweeps_trace[iter].
 */
int orc_ctpf_train(int64_t K, int64_t M, int64_t V, int64_t U, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts,
                   const int64_t *R_cumsum, const int64_t *readers, const int64_t *ratings, const double *hyp, double *alef,
                   double *alef_old, double *he, double *he_old, double *bet, double *bet_old, double *vav, double *vav_old,
                   double *gimel, double *gimel_old, double *zayin, double *zayin_old, double *dalet, double *dalet_old, double *het,
                   double *het_old, int iter, double tol, int viter, double vtol, int checkelbo, double *elbo_trace,
                   int64_t *sweeps_trace, int *iters_done, int nthreads)
{
    const double a = hyp[0], b = hyp[1], c = hyp[2], dd = hyp[3], e = hyp[4], f = hyp[5], g = hyp[6], h = hyp[7];
    if (nthreads < 1) nthreads = 1;
    int64_t KV = K * V, KU = K * (U > 0 ? U : 1), mxn = 1, mxr = 1;
    int all_empty = 1;
    for (int64_t d = 0; d < M; d++) {
        if (N_cumsum[d + 1] > N_cumsum[d]) all_empty = 0;
        if (N_cumsum[d + 1] - N_cumsum[d] > mxn) mxn = N_cumsum[d + 1] - N_cumsum[d];
        if (R_cumsum[d + 1] - R_cumsum[d] > mxr) mxr = R_cumsum[d + 1] - R_cumsum[d];
    }
    if (all_empty) iter = 0;
    memcpy(alef_old, alef, sizeof(double) * KV);
    memcpy(he_old, he, sizeof(double) * K * U);
    memcpy(bet_old, bet, sizeof(double) * K);
    memcpy(vav_old, vav, sizeof(double) * K);
    memcpy(dalet_old, dalet, sizeof(double) * K);
    memcpy(het_old, het, sizeof(double) * K);
    memcpy(gimel_old, gimel, sizeof(double) * K * M);
    memcpy(zayin_old, zayin, sizeof(double) * K * M);
    for (int k = 0; k <= iter; k++) elbo_trace[k] = NAN;
    int check = checkelbo > 0;
    double elbo = 0.0;
#define ELBO_NOW() orc_ctpf_elbo(K, M, V, U, N_cumsum, terms, counts, R_cumsum, readers, ratings, hyp, alef, alef_old, he, he_old, bet, \
                                 bet_old, vav, vav_old, gimel, gimel_old, zayin, zayin_old, dalet, dalet_old, het, het_old, nthreads)
    if (check && checkelbo <= iter) {
        elbo = ELBO_NOW();
        elbo_trace[0] = elbo;
    }
    double *alef_temp = (double *)calloc((size_t)KV * nthreads, sizeof(double));
    double *he_temp = (double *)calloc((size_t)KU * nthreads, sizeof(double));
    double *gsum = (double *)malloc(sizeof(double) * K), *zsum = (double *)malloc(sizeof(double) * K);
    double *alsum = (double *)malloc(sizeof(double) * K), *hesum = (double *)malloc(sizeof(double) * K);
    int k_done = 0;
    for (int k = 1; k <= iter; k++) {
        int64_t sweeps = 0;
#pragma omp parallel num_threads(nthreads) reduction(+ : sweeps)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
            double *at = alef_temp + (size_t)KV * tid, *ht = he_temp + (size_t)KU * tid;
            double *phi = (double *)malloc(sizeof(double) * K * mxn), *xi = (double *)malloc(sizeof(double) * 2 * K * mxr);
#pragma omp for schedule(dynamic, 16)
            for (int64_t d = 0; d < M; d++) {
                int64_t o = N_cumsum[d], Nd = N_cumsum[d + 1] - o, ro = R_cumsum[d], Rd = R_cumsum[d + 1] - ro;
                double *gm = gimel + K * d, *zy = zayin + K * d, *go = gimel_old + K * d, *zo = zayin_old + K * d;
                for (int v = 0; v < viter; v++) {                                            /* CTPF.jl:354-362 */
                    ctpf_xi(K, Rd, readers + ro, he, gm, zy, dalet, het, vav, xi);
                    ctpf_phi(K, Nd, terms + o, alef, gm, dalet, bet, phi);
                    for (int64_t i = 0; i < K; i++) { zo[i] = zy[i]; zy[i] = g; }            /* update_zayin!, CTPF.jl:318-323 */
                    for (int64_t r = 0; r < Rd; r++)
                        for (int64_t i = 0; i < K; i++) zy[i] += xi[2 * K * r + K + i] * (double)ratings[ro + r];
                    for (int64_t i = 0; i < K; i++) { go[i] = gm[i]; gm[i] = c; }            /* update_gimel!, CTPF.jl:309-314 */
                    for (int64_t n = 0; n < Nd; n++)
                        for (int64_t i = 0; i < K; i++) gm[i] += phi[K * n + i] * (double)counts[o + n];
                    for (int64_t r = 0; r < Rd; r++)
                        for (int64_t i = 0; i < K; i++) gm[i] += xi[2 * K * r + i] * (double)ratings[ro + r];
                    sweeps++;
                    double nrm = 0.0;
                    for (int64_t i = 0; i < K; i++) nrm += (gm[i] - go[i]) * (gm[i] - go[i]);
                    if (sqrt(nrm) < vtol) break;
                }
                for (int64_t r = 0; r < Rd; r++)                                             /* update_he!(d), CTPF.jl:274-277 */
                    for (int64_t i = 0; i < K; i++)
                        ht[K * readers[ro + r] + i] += (xi[2 * K * r + i] + xi[2 * K * r + K + i]) * (double)ratings[ro + r];
                for (int64_t n = 0; n < Nd; n++)                                             /* update_alef!(d), CTPF.jl:259-262 */
                    for (int64_t i = 0; i < K; i++) at[K * terms[o + n] + i] += phi[K * n + i] * (double)counts[o + n];
            }
            free(phi);
            free(xi);
        }
        for (int t = 1; t < nthreads; t++) {
            double *at = alef_temp + (size_t)KV * t, *ht = he_temp + (size_t)KU * t;
            for (int64_t q = 0; q < KV; q++) { alef_temp[q] += at[q]; at[q] = 0.0; }
            for (int64_t q = 0; q < K * U; q++) { he_temp[q] += ht[q]; ht[q] = 0.0; }
        }
        if (sweeps_trace) sweeps_trace[k - 1] = sweeps;
        /* CTPF.jl:366-371: he, alef, dalet, het, bet, vav (in this order) */
        memcpy(he_old, he, sizeof(double) * K * U);
        for (int64_t q = 0; q < K * U; q++) { he[q] = e + he_temp[q]; he_temp[q] = 0.0; }
        memcpy(alef_old, alef, sizeof(double) * KV);
        for (int64_t q = 0; q < KV; q++) { alef[q] = a + alef_temp[q]; alef_temp[q] = 0.0; }
        for (int64_t i = 0; i < K; i++) { alsum[i] = hesum[i] = gsum[i] = zsum[i] = 0.0; }
        for (int64_t j = 0; j < V; j++) for (int64_t i = 0; i < K; i++) alsum[i] += alef[K * j + i];
        for (int64_t u = 0; u < U; u++) for (int64_t i = 0; i < K; i++) hesum[i] += he[K * u + i];
        for (int64_t d = 0; d < M; d++) for (int64_t i = 0; i < K; i++) { gsum[i] += gimel[K * d + i]; zsum[i] += zayin[K * d + i]; }
        for (int64_t i = 0; i < K; i++) {
            dalet_old[i] = dalet[i];
            dalet[i] = dd + alsum[i] / bet[i] + hesum[i] / vav[i];          /* update_dalet!, CTPF.jl:295-298 */
            het_old[i] = het[i];
            het[i] = h + hesum[i] / vav[i];                                 /* update_het!, CTPF.jl:302-305 */
        }
        for (int64_t i = 0; i < K; i++) {
            bet_old[i] = bet[i];
            bet[i] = b + gsum[i] / dalet[i];                                /* update_bet!, CTPF.jl:281-284 */
            vav_old[i] = vav[i];
            vav[i] = f + gsum[i] / dalet[i] + zsum[i] / het[i];             /* update_vav!, CTPF.jl:288-291 */
        }
        k_done = k;
        if (check && (k % checkelbo == 0)) {
            double e2 = ELBO_NOW();
            double delta = e2 - elbo;
            elbo = e2;
            elbo_trace[k] = e2;
            if (delta < tol) break;
        }
    }
#undef ELBO_NOW
    free(alef_temp); free(he_temp); free(gsum); free(zsum); free(alsum); free(hesum);
    if (iters_done) *iters_done = k_done;
    return 0;
}
