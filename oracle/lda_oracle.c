/*
 * oracle/lda_oracle.c -- fp64 CPU restatement of the reference's CPU LDA (src/LDA.jl).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under topicmodelsvb.jl_b200/ may link, import or call
 * this file; it is the checker for tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or known-answer values
 * (v0.6/test/runtests.jl is empty) and Julia is not installed in this image, so this
 * restatement cannot be checked against the reference's own output.  It is cross-checked
 * against an independent NumPy/SciPy twin (oracle/numpy_twin.py) to <= 1e-12 relative.
 *
 * Array conventions follow the reference's host structs: dense matrices are Julia
 * column-major, i.e. beta[K*j + i] is topic i of term j; Elogtheta[K*d + i], gamma[K*d + i].
 * terms are 0-based here (modelutils.jl:371 subtracts 1 the same way).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "special.h"

#define ORC_EPS 0x1p-99 /* EPSILON = eps(1e-14) = 2^-99, utils.jl:3 */

static inline double orc_finite(double x) /* utils.jl:107 */
{
    double a = fabs(x);
    if (a > DBL_MAX) a = DBL_MAX;
    return copysign(a, x);
}

/* LDA.jl:150-154  update_phi!: phi = EPS + beta[:,terms] .* exp.(Elogtheta[d]); column-normalise. */
static void lda_update_phi(int64_t K, int64_t Nd, const int64_t *terms, const double *beta,
                           const double *Elogtheta_d, double *expEt, double *phi)
{
    for (int64_t i = 0; i < K; i++) expEt[i] = exp(Elogtheta_d[i]);
    for (int64_t n = 0; n < Nd; n++) {
        const double *b = beta + K * terms[n];
        double *p = phi + K * n;
        double s = 0.0;
        for (int64_t i = 0; i < K; i++) { p[i] = ORC_EPS + b[i] * expEt[i]; s += p[i]; }
        for (int64_t i = 0; i < K; i++) p[i] /= s;
    }
}

/* LDA.jl:143-146  update_gamma!: gamma[d] = EPS + (alpha + phi * counts). */
static void lda_update_gamma(int64_t K, int64_t Nd, const int64_t *counts, const double *alpha,
                             const double *phi, double *gamma_d)
{
    for (int64_t i = 0; i < K; i++) gamma_d[i] = 0.0;
    for (int64_t n = 0; n < Nd; n++) {
        const double *p = phi + K * n;
        double c = (double)counts[n];
        for (int64_t i = 0; i < K; i++) gamma_d[i] += p[i] * c;
    }
    for (int64_t i = 0; i < K; i++) gamma_d[i] = ORC_EPS + (alpha[i] + gamma_d[i]);
}

/* LDA.jl:136-139  update_Elogtheta!: old <- cur; cur = digamma.(gamma) .- digamma(sum(gamma)). */
static void lda_update_Elogtheta(int64_t K, const double *gamma_d, double *Elogtheta_d, double *Elogtheta_old_d)
{
    double g0 = 0.0;
    for (int64_t i = 0; i < K; i++) g0 += gamma_d[i];
    double dg0 = orc_digamma(g0);
    for (int64_t i = 0; i < K; i++) {
        Elogtheta_old_d[i] = Elogtheta_d[i];
        Elogtheta_d[i] = orc_digamma(gamma_d[i]) - dg0;
    }
}

/* LDA.jl:170-178  the per-document inner loop.  Returns the number of sweeps made.
 * On exit phi is the LAST phi computed (from the Elogtheta now held in Elogtheta_old_d),
 * which is what update_beta!(model, d) (LDA.jl:129-132) scatters. */
static int lda_doc_estep(int64_t K, int64_t Nd, const int64_t *terms, const int64_t *counts,
                         const double *beta, const double *alpha, double *Elogtheta_d,
                         double *Elogtheta_old_d, double *gamma_d, double *expEt, double *phi,
                         int viter, double vtol)
{
    int v = 0;
    for (v = 0; v < viter; v++) {
        lda_update_phi(K, Nd, terms, beta, Elogtheta_d, expEt, phi);
        lda_update_gamma(K, Nd, counts, alpha, phi, gamma_d);
        lda_update_Elogtheta(K, gamma_d, Elogtheta_d, Elogtheta_old_d);
        double nrm = 0.0;
        for (int64_t i = 0; i < K; i++) {
            double df = Elogtheta_d[i] - Elogtheta_old_d[i];
            nrm += df * df;
        }
        if (sqrt(nrm) < vtol) { v++; break; }
    }
    return v;
}

/* LDA.jl:50-93  one document's ELBO contribution with the lagged-phi reconstruction
 * (phi from beta_old and Elogtheta_old, LDA.jl:87-88), evaluated with the current
 * alpha, beta, gamma[d], Elogtheta[d]. */
static double lda_doc_elbo(int64_t K, int64_t Nd, const int64_t *terms, const int64_t *counts,
                           const double *alpha, double lg_alpha_term, const double *beta,
                           const double *beta_old, const double *Elogtheta_d,
                           const double *Elogtheta_old_d, const double *gamma_d, double *expEt,
                           double *phi)
{
    lda_update_phi(K, Nd, terms, beta_old, Elogtheta_old_d, expEt, phi);

    /* Elogptheta, LDA.jl:50-53 */
    double x = lg_alpha_term;
    for (int64_t i = 0; i < K; i++) x += (alpha[i] - 1.0) * Elogtheta_d[i];

    double elogpz = 0.0, elogpw = 0.0, negelogqz = 0.0;
    for (int64_t n = 0; n < Nd; n++) {
        const double *p = phi + K * n;
        const double *b = beta + K * terms[n];
        double c = (double)counts[n];
        double pz = 0.0, pw = 0.0, ent = 0.0;
        for (int64_t i = 0; i < K; i++) {
            pz += p[i] * Elogtheta_d[i];                   /* Elogpz, LDA.jl:57-60 */
            pw += p[i] * log(b[i] + ORC_EPS);             /* Elogpw, LDA.jl:64-67 (@boink) */
            if (p[i] > 0.0) ent -= p[i] * log(p[i]);      /* entropy(Categorical), LDA.jl:78 */
        }
        elogpz += c * pz;
        elogpw += c * pw;
        negelogqz += c * ent;
    }

    /* -Elogqtheta = entropy(Dirichlet(gamma[d])), LDA.jl:70-73 with the override at utils.jl:163-180 */
    double ent_dir = 0.0;
    if (K > 1) {
        double g0 = 0.0, lmnB = 0.0;
        for (int64_t i = 0; i < K; i++) { g0 += gamma_d[i]; lmnB += orc_lgamma(gamma_d[i]); }
        lmnB -= orc_lgamma(g0);
        ent_dir = lmnB + (g0 - (double)K) * orc_digamma(g0);
        for (int64_t i = 0; i < K; i++) ent_dir -= (gamma_d[i] - 1.0) * orc_digamma(gamma_d[i]);
    }
    return x + elogpz + elogpw + ent_dir + negelogqz;
}

static int64_t max_doc_len(int64_t M, const int64_t *N_cumsum)
{
    int64_t mx = 1;
    for (int64_t d = 0; d < M; d++)
        if (N_cumsum[d + 1] - N_cumsum[d] > mx) mx = N_cumsum[d + 1] - N_cumsum[d];
    return mx;
}

/* LDA.jl:83-93  update_elbo!.  Documents are independent; with nthreads > 1 the per-document
 * terms are summed per thread and then across threads (fp64 re-association only). */
double orc_lda_elbo(int64_t K, int64_t M, int64_t V, const int64_t *N_cumsum, const int64_t *terms,
                    const int64_t *counts, const double *alpha, const double *beta,
                    const double *beta_old, const double *Elogtheta, const double *Elogtheta_old,
                    const double *gamma, int nthreads)
{
    (void)V;
    double a0 = 0.0, sl = 0.0;
    for (int64_t i = 0; i < K; i++) { a0 += alpha[i]; sl += orc_lgamma(alpha[i]); }
    double lg_alpha_term = orc_finite(orc_lgamma(a0)) - orc_finite(sl);
    int64_t mx = max_doc_len(M, N_cumsum);
    double elbo = 0.0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads) reduction(+ : elbo)
    {
        double *phi = (double *)malloc(sizeof(double) * K * mx);
        double *expEt = (double *)malloc(sizeof(double) * K);
#pragma omp for schedule(dynamic, 64)
        for (int64_t d = 0; d < M; d++) {
            int64_t o = N_cumsum[d], Nd = N_cumsum[d + 1] - o;
            elbo += lda_doc_elbo(K, Nd, terms + o, counts + o, alpha, lg_alpha_term, beta, beta_old,
                                 Elogtheta + K * d, Elogtheta_old + K * d, gamma + K * d, expEt, phi);
        }
        free(phi);
        free(expEt);
    }
    return elbo;
}

/* LDA.jl:97-118  update_alpha!: interior-point Newton with log barrier nu (halved each step),
 * linear-time Hessian inverse (diagonal + rank one), back-tracking to keep alpha >= 0. */
int orc_lda_update_alpha(int64_t K, int64_t M, double *alpha, const double *Elogtheta_sum, int niter,
                         double ntol)
{
    double *grad = (double *)malloc(sizeof(double) * K);
    double *hinv = (double *)malloc(sizeof(double) * K);
    double *p = (double *)malloc(sizeof(double) * K);
    double nu = (double)K, Md = (double)M;
    int it = 0;
    for (it = 0; it < niter; it++) {
        double rho = 1.0, a0 = 0.0;
        for (int64_t i = 0; i < K; i++) a0 += alpha[i];
        double dg0 = orc_digamma(a0);
        double gh = 0.0, hs = 0.0, gn = 0.0;
        for (int64_t i = 0; i < K; i++) {
            grad[i] = nu / alpha[i] + Md * (dg0 - orc_digamma(alpha[i])) + Elogtheta_sum[i];
            hinv[i] = -1.0 / (Md * orc_trigamma(alpha[i]) + nu / (alpha[i] * alpha[i]));
            gh += grad[i] * hinv[i];
            hs += hinv[i];
            gn += grad[i] * grad[i];
        }
        double z = gh / (1.0 / (Md * orc_trigamma(a0)) + hs);
        for (int64_t i = 0; i < K; i++) p[i] = (grad[i] - z) * hinv[i];
        for (;;) {
            double mn = INFINITY;
            for (int64_t i = 0; i < K; i++) {
                double t = alpha[i] - rho * p[i];
                if (t < mn) mn = t;
            }
            if (!(mn < 0.0)) break;
            rho *= 0.5;
        }
        /* @finite alpha -= rho*p  ==>  sign(alpha) * min(|alpha - rho p|, floatmax), macros.jl:52-54 */
        for (int64_t i = 0; i < K; i++) {
            double t = fabs(alpha[i] - rho * p[i]);
            if (t > DBL_MAX) t = DBL_MAX;
            alpha[i] = copysign(t, alpha[i]);
        }
        if ((rho * sqrt(gn) < ntol) && (nu / (double)K < ntol)) { it++; break; }
        nu *= 0.5;
    }
    for (int64_t i = 0; i < K; i++) alpha[i] += ORC_EPS; /* @positive model.alpha */
    free(grad);
    free(hinv);
    free(p);
    return it;
}

/*
 * LDA.jl:161-191  train!.
 *
 * In/out (caller-allocated): alpha[K], beta[K*V], Elogtheta[K*M], gamma[K*M].
 * Out: beta_old[K*V], Elogtheta_old[K*M] (the lagged copies the struct keeps),
 *      elbo_trace[iter+1] (slot 0 = the initial update_elbo!, slot k = ELBO after iteration k
 *      when k % checkelbo == 0, NaN otherwise), sweeps_trace[iter] (total inner sweeps of
 *      iteration k, diagnostic), *iters_done.
 * checkelbo <= 0 means Inf (never evaluate).  nthreads == 1 follows the reference's exact
 * sequential order; nthreads > 1 distributes documents over OpenMP threads with per-thread
 * beta_temp copies that are summed afterwards (fp64 re-association only) -- used for the
 * all-cores CPU baseline.
 */
int orc_lda_train(int64_t K, int64_t M, int64_t V, const int64_t *N_cumsum, const int64_t *terms,
                  const int64_t *counts, double *alpha, double *beta, double *beta_old,
                  double *Elogtheta, double *Elogtheta_old, double *gamma, int iter, double tol,
                  int niter, double ntol, int viter, double vtol, int checkelbo, double *elbo_trace,
                  int64_t *sweeps_trace, int *iters_done, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    int64_t KV = K * V, mx = max_doc_len(M, N_cumsum);
    int all_empty = 1;
    for (int64_t d = 0; d < M; d++)
        if (N_cumsum[d + 1] > N_cumsum[d]) { all_empty = 0; break; }
    if (all_empty) iter = 0; /* LDA.jl:166 */

    memcpy(beta_old, beta, sizeof(double) * KV);
    memcpy(Elogtheta_old, Elogtheta, sizeof(double) * K * M);
    for (int k = 0; k <= iter; k++) elbo_trace[k] = NAN;

    double elbo = 0.0;
    int check = (checkelbo > 0);
    if (check && checkelbo <= iter) { /* LDA.jl:167 */
        elbo = orc_lda_elbo(K, M, V, N_cumsum, terms, counts, alpha, beta, beta_old, Elogtheta,
                            Elogtheta_old, gamma, nthreads);
        elbo_trace[0] = elbo;
    }

    double *beta_temp = (double *)calloc((size_t)KV * (size_t)nthreads, sizeof(double));
    double *Esum = (double *)malloc(sizeof(double) * K);
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
            double *bt = beta_temp + (size_t)KV * tid;
            double *phi = (double *)malloc(sizeof(double) * K * mx);
            double *expEt = (double *)malloc(sizeof(double) * K);
#pragma omp for schedule(dynamic, 64)
            for (int64_t d = 0; d < M; d++) {
                int64_t o = N_cumsum[d], Nd = N_cumsum[d + 1] - o;
                sweeps += lda_doc_estep(K, Nd, terms + o, counts + o, beta, alpha, Elogtheta + K * d,
                                        Elogtheta_old + K * d, gamma + K * d, expEt, phi, viter, vtol);
                /* LDA.jl:129-132  update_beta!(model, d): beta_temp[:,terms] += phi .* counts' */
                for (int64_t n = 0; n < Nd; n++) {
                    double *b = bt + K * terms[o + n];
                    const double *p = phi + K * n;
                    double c = (double)counts[o + n];
                    for (int64_t i = 0; i < K; i++) b[i] += p[i] * c;
                }
            }
            free(phi);
            free(expEt);
        }
        for (int t = 1; t < nthreads; t++) {
            double *bt = beta_temp + (size_t)KV * t;
            for (int64_t q = 0; q < KV; q++) { beta_temp[q] += bt[q]; bt[q] = 0.0; }
        }
        if (sweeps_trace) sweeps_trace[k - 1] = sweeps;

        /* LDA.jl:121-125  update_beta!(model): beta_old <- beta; beta = beta_temp ./ rowsum; beta_temp <- 0 */
        memcpy(beta_old, beta, sizeof(double) * KV);
        for (int64_t i = 0; i < K; i++) {
            double rs = 0.0;
            for (int64_t j = 0; j < V; j++) rs += beta_temp[K * j + i];
            for (int64_t j = 0; j < V; j++) beta[K * j + i] = beta_temp[K * j + i] / rs;
        }
        memset(beta_temp, 0, sizeof(double) * KV);

        /* LDA.jl:98  Elogtheta_sum = sum over documents */
        for (int64_t i = 0; i < K; i++) Esum[i] = 0.0;
        for (int64_t d = 0; d < M; d++)
            for (int64_t i = 0; i < K; i++) Esum[i] += Elogtheta[K * d + i];
        orc_lda_update_alpha(K, M, alpha, Esum, niter, ntol);

        k_done = k;
        /* modelutils.jl:574-585  check_elbo! */
        if (check && (k % checkelbo == 0)) {
            double e2 = orc_lda_elbo(K, M, V, N_cumsum, terms, counts, alpha, beta, beta_old, Elogtheta,
                                     Elogtheta_old, gamma, nthreads);
            double delta = e2 - elbo;
            elbo = e2;
            elbo_trace[k] = e2;
            if (delta < tol) break;
        }
    }
    free(beta_temp);
    free(Esum);
    if (iters_done) *iters_done = k_done;
    return 0;
}

/* E-step only (LDA.jl:170-178 for every document, no M-step): the unit bench.py's cpu_baseline
 * times, and what predict (modelutils.jl:831-855) runs.  stats[K*V] receives the scatter. */
int64_t orc_lda_estep(int64_t K, int64_t M, int64_t V, const int64_t *N_cumsum, const int64_t *terms,
                      const int64_t *counts, const double *alpha, const double *beta, double *Elogtheta,
                      double *Elogtheta_old, double *gamma, double *stats, int viter, double vtol,
                      int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    int64_t KV = K * V, mx = max_doc_len(M, N_cumsum), sweeps = 0;
    double *extra = nthreads > 1 ? (double *)calloc((size_t)KV * (size_t)(nthreads - 1), sizeof(double)) : NULL;
#pragma omp parallel num_threads(nthreads) reduction(+ : sweeps)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        double *bt = tid == 0 ? stats : extra + (size_t)KV * (tid - 1);
        double *phi = (double *)malloc(sizeof(double) * K * mx);
        double *expEt = (double *)malloc(sizeof(double) * K);
#pragma omp for schedule(dynamic, 64)
        for (int64_t d = 0; d < M; d++) {
            int64_t o = N_cumsum[d], Nd = N_cumsum[d + 1] - o;
            sweeps += lda_doc_estep(K, Nd, terms + o, counts + o, beta, alpha, Elogtheta + K * d,
                                    Elogtheta_old + K * d, gamma + K * d, expEt, phi, viter, vtol);
            if (stats)
                for (int64_t n = 0; n < Nd; n++) {
                    double *b = bt + K * terms[o + n];
                    const double *p = phi + K * n;
                    double c = (double)counts[o + n];
                    for (int64_t i = 0; i < K; i++) b[i] += p[i] * c;
                }
        }
        free(phi);
        free(expEt);
    }
    if (stats)
        for (int t = 1; t < nthreads; t++) {
            double *bt = extra + (size_t)KV * (t - 1);
            for (int64_t q = 0; q < KV; q++) stats[q] += bt[q];
        }
    free(extra);
    return sweeps;
}

/* The last phi of every document (K x sum(N), column-major per token), reconstructed the way
 * update_elbo! does (LDA.jl:87-88); used to check tmvb_lda_materialize_phi. */
void orc_lda_phi(int64_t K, int64_t M, const int64_t *N_cumsum, const int64_t *terms,
                 const double *beta_old, const double *Elogtheta_old, double *phi)
{
    double *expEt = (double *)malloc(sizeof(double) * K);
    for (int64_t d = 0; d < M; d++) {
        int64_t o = N_cumsum[d], Nd = N_cumsum[d + 1] - o;
        lda_update_phi(K, Nd, terms + o, beta_old, Elogtheta_old + K * d, expEt, phi + K * o);
    }
    free(expEt);
}

double orc_digamma_export(double x) { return orc_digamma(x); }
double orc_trigamma_export(double x) { return orc_trigamma(x); }
double orc_lgamma_export(double x) { return orc_lgamma(x); }
