/*
 * oracle/flda_oracle.c -- fp64 CPU restatement of the reference's filtered LDA (src/fLDA.jl).
 *
 * TEST INFRASTRUCTURE ONLY (see lda_oracle.c): the checker of the gpufLDA path, never linked,
 * imported or called by anything under topicmodelsvb.jl_b200/.
 *
 * PARITY UNPINNED, like the other oracles: the reference ships no golden vectors and Julia is
 * not installed here.  Cross-checked against the literal NumPy transcription in
 * oracle/numpy_twin.py (FLDATwin) to <= 1e-12 relative (tests/test_oracle_cpu.py).
 *
 * Layout as in lda_oracle.c: beta[K*j + i] is topic i of term j (Julia's K x V column-major),
 * Elogtheta[K*d + i], gamma[K*d + i]; tau / tau_old are flat over the CSR tokens (tau[d][n] of the
 * reference at N_cumsum[d] + n); terms are 0-based.
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

int orc_lda_update_alpha(int64_t K, int64_t M, double *alpha, const double *Elogtheta_sum, int niter, double ntol); /* fLDA.jl:122-146 == LDA.jl:97-118 */

static inline double flda_finite(double x) /* utils.jl:107 */
{
    double a = fabs(x);
    if (a > DBL_MAX) a = DBL_MAX;
    return copysign(a, x);
}

/* fLDA.jl:198-201  update_phi!: additive_logistic(tau' .* log.(@boink beta[:,terms]) .+ Elogtheta[d], dims=1)
 * (utils.jl:114-121: subtract the column maximum, exp, normalise the column). */
static void flda_update_phi(int64_t K, int64_t Nd, const int64_t *terms, const double *beta, const double *tau_d,
                            const double *Elogtheta_d, double *phi)
{
    for (int64_t n = 0; n < Nd; n++) {
        const double *b = beta + K * terms[n];
        double *p = phi + K * n;
        double mx = -INFINITY, s = 0.0;
        for (int64_t i = 0; i < K; i++) {
            p[i] = tau_d[n] * log(b[i] + ORC_EPS) + Elogtheta_d[i];
            if (p[i] > mx) mx = p[i];
        }
        for (int64_t i = 0; i < K; i++) { p[i] = exp(p[i] - mx); s += p[i]; }
        for (int64_t i = 0; i < K; i++) p[i] /= s;
    }
}

/* fLDA.jl:189-194  update_tau!: tau_old <- tau; tau = eta ./ (@boink eta .+ (1 - eta) * (kappa[terms] .* prod(beta[:,terms].^-phi, dims=1))) */
static void flda_update_tau(int64_t K, int64_t Nd, const int64_t *terms, const double *beta, const double *kappa, double eta,
                            const double *phi, double *tau_d, double *tau_old_d)
{
    for (int64_t n = 0; n < Nd; n++) {
        const double *b = beta + K * terms[n];
        const double *p = phi + K * n;
        double pr = 1.0;
        for (int64_t i = 0; i < K; i++) pr *= pow(b[i], -p[i]);
        tau_old_d[n] = tau_d[n];
        tau_d[n] = eta / ((eta + (1.0 - eta) * (kappa[terms[n]] * pr)) + ORC_EPS);
    }
}

/* fLDA.jl:182-185  update_gamma!: @positive gamma[d] = alpha + phi * counts */
static void flda_update_gamma(int64_t K, int64_t Nd, const int64_t *counts, const double *alpha, const double *phi, double *gamma_d)
{
    for (int64_t i = 0; i < K; i++) gamma_d[i] = 0.0;
    for (int64_t n = 0; n < Nd; n++) {
        const double *p = phi + K * n;
        double c = (double)counts[n];
        for (int64_t i = 0; i < K; i++) gamma_d[i] += p[i] * c;
    }
    for (int64_t i = 0; i < K; i++) gamma_d[i] = ORC_EPS + (alpha[i] + gamma_d[i]);
}

/* fLDA.jl:175-178  update_Elogtheta! */
static void flda_update_Elogtheta(int64_t K, const double *gamma_d, double *Elogtheta_d, double *Elogtheta_old_d)
{
    double g0 = 0.0;
    for (int64_t i = 0; i < K; i++) g0 += gamma_d[i];
    double dg0 = orc_digamma(g0);
    for (int64_t i = 0; i < K; i++) {
        Elogtheta_old_d[i] = Elogtheta_d[i];
        Elogtheta_d[i] = orc_digamma(gamma_d[i]) - dg0;
    }
}

/* fLDA.jl:223-233  the per-document inner loop; returns the number of sweeps.  On exit phi is the last phi. */
static int flda_doc_estep(int64_t K, int64_t Nd, const int64_t *terms, const int64_t *counts, const double *beta, const double *kappa,
                          double eta, const double *alpha, double *Elogtheta_d, double *Elogtheta_old_d, double *gamma_d, double *tau_d,
                          double *tau_old_d, double *phi, int viter, double vtol)
{
    int v = 0;
    for (v = 0; v < viter; v++) {
        flda_update_phi(K, Nd, terms, beta, tau_d, Elogtheta_d, phi);
        flda_update_tau(K, Nd, terms, beta, kappa, eta, phi, tau_d, tau_old_d);
        flda_update_gamma(K, Nd, counts, alpha, phi, gamma_d);
        flda_update_Elogtheta(K, gamma_d, Elogtheta_d, Elogtheta_old_d);
        double nrm = 0.0;
        for (int64_t i = 0; i < K; i++) {
            double df = Elogtheta_d[i] - Elogtheta_old_d[i];
            nrm += df * df;
        }
        if (sqrt(nrm) < vtol) { v++; break; }
    }
    return v;
}

/* entropy(Bernoulli(p)) of Distributions.jl (univariate/discrete/bernoulli.jl): 0 at p in {0, 1} */
static double bernoulli_entropy(double p)
{
    double p0 = 1.0 - p;
    if (p0 == 0.0 || p0 == 1.0) return 0.0;
    return -(p0 * log(p0) + p * log(p));
}

/* fLDA.jl:62-117  one document's ELBO terms with the lagged phi of update_elbo! (fLDA.jl:105-108) */
static double flda_doc_elbo(int64_t K, int64_t Nd, int64_t Cd, const int64_t *terms, const int64_t *counts, const double *alpha,
                            double lg_alpha_term, double eta, const double *kappa, const double *beta, const double *beta_old,
                            const double *Elogtheta_d, const double *Elogtheta_old_d, const double *gamma_d, const double *tau_d,
                            const double *tau_old_d, double *phi)
{
    flda_update_phi(K, Nd, terms, beta_old, tau_old_d, Elogtheta_old_d, phi);

    double x = lg_alpha_term; /* Elogptheta, fLDA.jl:62-65 */
    for (int64_t i = 0; i < K; i++) x += (alpha[i] - 1.0) * Elogtheta_d[i];

    double tc = 0.0; /* Elogpc, fLDA.jl:68-72 */
    for (int64_t n = 0; n < Nd; n++) tc += tau_d[n] * (double)counts[n];
    double elogpc = log(pow(eta, tc) * pow(1.0 - eta, (double)Cd - tc) + ORC_EPS);

    double elogpz = 0.0, elogpw = 0.0, negelogqc = 0.0, negelogqz = 0.0;
    for (int64_t n = 0; n < Nd; n++) {
        const double *p = phi + K * n;
        const double *b = beta + K * terms[n];
        double c = (double)counts[n];
        double pz = 0.0, pw = 0.0, ent = 0.0;
        for (int64_t i = 0; i < K; i++) {
            pz += p[i] * Elogtheta_d[i];                  /* Elogpz, fLDA.jl:75-79 */
            pw += p[i] * log(b[i] + ORC_EPS);             /* Elogpw, fLDA.jl:82-86, first sum */
            if (p[i] > 0.0) ent -= p[i] * log(p[i]);      /* entropy(Categorical), fLDA.jl:103 */
        }
        elogpz += c * pz;
        elogpw += c * tau_d[n] * pw + c * (1.0 - tau_d[n]) * log(kappa[terms[n]] + ORC_EPS);
        negelogqc += c * bernoulli_entropy(tau_d[n]);     /* fLDA.jl:96 */
        negelogqz += c * ent;
    }

    double ent_dir = 0.0; /* -Elogqtheta, fLDA.jl:89-92 with utils.jl:163-180 */
    if (K > 1) {
        double g0 = 0.0, lmnB = 0.0;
        for (int64_t i = 0; i < K; i++) { g0 += gamma_d[i]; lmnB += orc_lgamma(gamma_d[i]); }
        lmnB -= orc_lgamma(g0);
        ent_dir = lmnB + (g0 - (double)K) * orc_digamma(g0);
        for (int64_t i = 0; i < K; i++) ent_dir -= (gamma_d[i] - 1.0) * orc_digamma(gamma_d[i]);
    }
    return x + elogpc + elogpz + elogpw + ent_dir + negelogqc + negelogqz;
}

static int64_t flda_max_doc_len(int64_t M, const int64_t *N_cumsum)
{
    int64_t mx = 1;
    for (int64_t d = 0; d < M; d++)
        if (N_cumsum[d + 1] - N_cumsum[d] > mx) mx = N_cumsum[d + 1] - N_cumsum[d];
    return mx;
}

/* fLDA.jl:105-117  update_elbo! */
double orc_flda_elbo(int64_t K, int64_t M, int64_t V, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts, double eta,
                     const double *alpha, const double *kappa, const double *beta, const double *beta_old, const double *Elogtheta,
                     const double *Elogtheta_old, const double *gamma, const double *tau, const double *tau_old, int nthreads)
{
    (void)V;
    double a0 = 0.0, sl = 0.0;
    for (int64_t i = 0; i < K; i++) { a0 += alpha[i]; sl += orc_lgamma(alpha[i]); }
    double lg_alpha_term = flda_finite(orc_lgamma(a0)) - flda_finite(sl);
    int64_t mx = flda_max_doc_len(M, N_cumsum);
    double elbo = 0.0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads) reduction(+ : elbo)
    {
        double *phi = (double *)malloc(sizeof(double) * K * mx);
#pragma omp for schedule(dynamic, 64)
        for (int64_t d = 0; d < M; d++) {
            int64_t o = N_cumsum[d], Nd = N_cumsum[d + 1] - o, Cd = 0;
            for (int64_t n = 0; n < Nd; n++) Cd += counts[o + n];
            elbo += flda_doc_elbo(K, Nd, Cd, terms + o, counts + o, alpha, lg_alpha_term, eta, kappa, beta, beta_old, Elogtheta + K * d,
                                  Elogtheta_old + K * d, gamma + K * d, tau + o, tau_old + o, phi);
        }
        free(phi);
    }
    return elbo;
}

/*
 * fLDA.jl:214-247  train!.
 * In/out: eta[1], alpha[K], kappa[V], beta[K*V], Elogtheta[K*M], gamma[K*M], tau[nnz].
 * Out: kappa_old[V], beta_old[K*V], Elogtheta_old[K*M], tau_old[nnz], elbo_trace[iter+1] (slot 0 = the initial update_elbo!,
 *      NaN where not evaluated), sweeps_trace[iter], *iters_done.
 * nthreads == 1 follows the reference's sequential order; more threads split the documents (per-thread beta_temp / kappa_temp,
 * fp64 re-association only).
 */
int orc_flda_train(int64_t K, int64_t M, int64_t V, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts, double *eta,
                   double *alpha, double *kappa, double *kappa_old, double *beta, double *beta_old, double *Elogtheta, double *Elogtheta_old,
                   double *gamma, double *tau, double *tau_old, int iter, double tol, int niter, double ntol, int viter, double vtol,
                   int checkelbo, double *elbo_trace, int64_t *sweeps_trace, int *iters_done, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    const int64_t KV = K * V, nnz = N_cumsum[M], mx = flda_max_doc_len(M, N_cumsum);
    int all_empty = 1;
    for (int64_t d = 0; d < M; d++)
        if (N_cumsum[d + 1] > N_cumsum[d]) { all_empty = 0; break; }
    if (all_empty) iter = 0; /* fLDA.jl:219 */
    memcpy(beta_old, beta, sizeof(double) * KV);
    memcpy(kappa_old, kappa, sizeof(double) * V);
    memcpy(Elogtheta_old, Elogtheta, sizeof(double) * K * M);
    memcpy(tau_old, tau, sizeof(double) * nnz);
    for (int k = 0; k <= iter; k++) elbo_trace[k] = NAN;
    double Ctot = 0.0;
    for (int64_t q = 0; q < nnz; q++) Ctot += (double)counts[q];

    double elbo = 0.0;
    int check = (checkelbo > 0);
    if (check && checkelbo <= iter) { /* fLDA.jl:220 */
        elbo = orc_flda_elbo(K, M, V, N_cumsum, terms, counts, *eta, alpha, kappa, beta, beta_old, Elogtheta, Elogtheta_old, gamma, tau,
                             tau_old, nthreads);
        elbo_trace[0] = elbo;
    }
    double *beta_temp = (double *)calloc((size_t)(KV + V) * (size_t)nthreads, sizeof(double));
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
            double *bt = beta_temp + (size_t)(KV + V) * tid, *kt = bt + KV;
            double *phi = (double *)malloc(sizeof(double) * K * mx);
#pragma omp for schedule(dynamic, 64)
            for (int64_t d = 0; d < M; d++) {
                int64_t o = N_cumsum[d], Nd = N_cumsum[d + 1] - o;
                sweeps += flda_doc_estep(K, Nd, terms + o, counts + o, beta, kappa, *eta, alpha, Elogtheta + K * d, Elogtheta_old + K * d,
                                         gamma + K * d, tau + o, tau_old + o, phi, viter, vtol);
                for (int64_t n = 0; n < Nd; n++) {
                    double *b = bt + K * terms[o + n];
                    const double *p = phi + K * n;
                    double c = (double)counts[o + n], t = tau[o + n];
                    for (int64_t i = 0; i < K; i++) b[i] += p[i] * (t * c); /* update_beta!(model, d), fLDA.jl:168-171 */
                    kt[terms[o + n]] += (1.0 - t) * c;                      /* update_kappa!(model, d), fLDA.jl:156-159 */
                }
            }
            free(phi);
        }
        for (int t = 1; t < nthreads; t++) {
            double *bt = beta_temp + (size_t)(KV + V) * t;
            for (int64_t q = 0; q < KV + V; q++) { beta_temp[q] += bt[q]; bt[q] = 0.0; }
        }
        if (sweeps_trace) sweeps_trace[k - 1] = sweeps;

        memcpy(beta_old, beta, sizeof(double) * KV); /* update_beta!(model), fLDA.jl:161-165 */
        for (int64_t i = 0; i < K; i++) {
            double rs = 0.0;
            for (int64_t j = 0; j < V; j++) rs += beta_temp[K * j + i];
            for (int64_t j = 0; j < V; j++) beta[K * j + i] = beta_temp[K * j + i] / rs;
        }
        memcpy(kappa_old, kappa, sizeof(double) * V); /* update_kappa!(model), fLDA.jl:149-153 */
        {
            double ks = 0.0;
            for (int64_t j = 0; j < V; j++) ks += beta_temp[KV + j];
            for (int64_t j = 0; j < V; j++) kappa[j] = beta_temp[KV + j] / ks;
        }
        memset(beta_temp, 0, sizeof(double) * (KV + V));

        for (int64_t i = 0; i < K; i++) Esum[i] = 0.0; /* update_alpha!, fLDA.jl:122-146 */
        for (int64_t d = 0; d < M; d++)
            for (int64_t i = 0; i < K; i++) Esum[i] += Elogtheta[K * d + i];
        orc_lda_update_alpha(K, M, alpha, Esum, niter, ntol);

        {   /* update_eta!, fLDA.jl:119-121 */
            double tc = 0.0;
            for (int64_t d = 0; d < M; d++) {
                double a = 0.0;
                for (int64_t q = N_cumsum[d]; q < N_cumsum[d + 1]; q++) a += tau[q] * (double)counts[q];
                tc += a;
            }
            *eta = tc / Ctot;
        }
        k_done = k;
        if (check && (k % checkelbo == 0)) { /* check_elbo!, modelutils.jl:574-585 */
            double e2 = orc_flda_elbo(K, M, V, N_cumsum, terms, counts, *eta, alpha, kappa, beta, beta_old, Elogtheta, Elogtheta_old, gamma,
                                      tau, tau_old, nthreads);
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
