/*
 * oracle/ctm_oracle.c -- fp64 CPU restatement of the reference's CPU correlated topic model (src/CTM.jl).
 *
 * TEST INFRASTRUCTURE ONLY (see lda_oracle.c).  PARITY UNPINNED: the reference has no golden vectors
 * and cannot run here; cross-checked against oracle/numpy_twin.py (CTMTwin) to <= 1e-10 relative.
 *
 * Layout: beta[K*j + i] (K x V column-major), lambda[K*d + i], vsq[K*d + i], logzeta[d],
 * sigma / invsigma K x K (symmetric, either order).  terms are 0-based.
 *
 * Third-party arithmetic restated (LinearAlgebra stdlib / Distributions.jl 0.23, not vendored):
 *   `Symmetric \ v` and `inv`, `logdet` -> Cholesky (the matrices are SPD, check_model modelutils.jl:116-119);
 *   entropy(MvNormal(m, S)) = (k (log 2pi + 1) + logdet S) / 2;  logsumexp = max-shifted log-sum-exp.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_EPS 0x1p-99 /* utils.jl:3 */

/* in-place lower Cholesky of the K x K SPD matrix A (row-major, symmetric); returns 0 on success */
static int chol(int64_t K, double *A)
{
    for (int64_t j = 0; j < K; j++) {
        double s = A[j * K + j];
        for (int64_t k = 0; k < j; k++) s -= A[j * K + k] * A[j * K + k];
        if (!(s > 0.0)) return -1;
        double l = sqrt(s);
        A[j * K + j] = l;
        for (int64_t i = j + 1; i < K; i++) {
            double t = A[i * K + j];
            for (int64_t k = 0; k < j; k++) t -= A[i * K + k] * A[j * K + k];
            A[i * K + j] = t / l;
        }
    }
    return 0;
}

static void chol_solve(int64_t K, const double *L, double *b)
{
    for (int64_t i = 0; i < K; i++) {
        double t = b[i];
        for (int64_t k = 0; k < i; k++) t -= L[i * K + k] * b[k];
        b[i] = t / L[i * K + i];
    }
    for (int64_t i = K - 1; i >= 0; i--) {
        double t = b[i];
        for (int64_t k = i + 1; k < K; k++) t -= L[k * K + i] * b[k];
        b[i] = t / L[i * K + i];
    }
}

/* inv(A) and logdet(A) of an SPD matrix (CTM.jl:57,110) */
#ifdef ORC_CTM_HELPERS_ONLY
int orc_spd_inv_logdet(int64_t K, const double *A, double *Ainv, double *logdet);
#else
int orc_spd_inv_logdet(int64_t K, const double *A, double *Ainv, double *logdet)
{
    double *L = (double *)malloc(sizeof(double) * K * K);
    double *e = (double *)malloc(sizeof(double) * K);
    memcpy(L, A, sizeof(double) * K * K);
    if (chol(K, L)) { free(L); free(e); return -1; }
    double ld = 0.0;
    for (int64_t i = 0; i < K; i++) ld += 2.0 * log(L[i * K + i]);
    if (logdet) *logdet = ld;
    if (Ainv)
        for (int64_t j = 0; j < K; j++) {
            for (int64_t i = 0; i < K; i++) e[i] = (i == j) ? 1.0 : 0.0;
            chol_solve(K, L, e);
            for (int64_t i = 0; i < K; i++) Ainv[i * K + j] = e[i];
        }
    free(L);
    free(e);
    return 0;
}

#endif /* ORC_CTM_HELPERS_ONLY */

/* CTM.jl:175-178  update_phi!: phi = additive_logistic(log.(beta[:,terms]) .+ lambda[d], dims=1) (utils.jl:114-123) */
static void ctm_update_phi(int64_t K, int64_t Nd, const int64_t *terms, const double *beta, const double *lambda_d, double *phi)
{
    for (int64_t n = 0; n < Nd; n++) {
        const double *b = beta + K * terms[n];
        double *p = phi + K * n;
        double mx = -INFINITY, s = 0.0;
        for (int64_t i = 0; i < K; i++) { p[i] = log(b[i]) + lambda_d[i]; if (p[i] > mx) mx = p[i]; }
        for (int64_t i = 0; i < K; i++) { p[i] = exp(p[i] - mx); s += p[i]; }
        for (int64_t i = 0; i < K; i++) p[i] /= s;
    }
}

/* CTM.jl:169-171  update_logzeta!: logsumexp(lambda + vsq / 2) */
static double ctm_logzeta(int64_t K, const double *lambda_d, const double *vsq_d)
{
    double mx = -INFINITY, s = 0.0;
    for (int64_t i = 0; i < K; i++) { double x = lambda_d[i] + 0.5 * vsq_d[i]; if (x > mx) mx = x; }
    for (int64_t i = 0; i < K; i++) s += exp(lambda_d[i] + 0.5 * vsq_d[i] - mx);
    return mx + log(s);
}

/* CTM.jl:146-165  update_vsq!: per-coordinate Newton with back-tracking; then vsq .+= EPSILON */
static void ctm_update_vsq(int64_t K, double Cd, const double *invsigma, const double *lambda_d, double *vsq_d, double logzeta_d,
                           int niter, double ntol)
{
    for (int64_t i = 0; i < K; i++) {
        for (int it = 0; it < niter; it++) {
            double rho = 1.0;
            double ex = Cd * exp(lambda_d[i] + 0.5 * vsq_d[i] - logzeta_d);
            double grad = -0.5 * (invsigma[i * K + i] + ex - 1.0 / vsq_d[i]);
            double invhess = -1.0 / (0.25 * ex + 0.5 / (vsq_d[i] * vsq_d[i]));
            double p = invhess * grad;
            while (vsq_d[i] - rho * p <= 0.0) rho *= 0.5;
            vsq_d[i] -= rho * p;
            if (rho * fabs(grad) < ntol) break;
        }
    }
    for (int64_t i = 0; i < K; i++) vsq_d[i] += ORC_EPS;
}

/* CTM.jl:129-142  update_lambda!: Newton; lambda += (invsigma + C diag(w)) \ grad; stop when ||grad|| < ntol */
static void ctm_update_lambda(int64_t K, int64_t Nd, const int64_t *counts, double Cd, const double *mu, const double *invsigma,
                              const double *phi, double *lambda_d, double *lambda_old_d, const double *vsq_d, double logzeta_d,
                              int niter, double ntol, double *phic, double *grad, double *H)
{
    for (int64_t i = 0; i < K; i++) { lambda_old_d[i] = lambda_d[i]; phic[i] = 0.0; }
    for (int64_t n = 0; n < Nd; n++) {
        const double *p = phi + K * n;
        double c = (double)counts[n];
        for (int64_t i = 0; i < K; i++) phic[i] += p[i] * c;
    }
    for (int it = 0; it < niter; it++) {
        double gn = 0.0;
        for (int64_t i = 0; i < K; i++) {
            double a = 0.0;
            for (int64_t j = 0; j < K; j++) a += invsigma[i * K + j] * (mu[j] - lambda_d[j]);
            double w = Cd * exp(lambda_d[i] + 0.5 * vsq_d[i] - logzeta_d);
            grad[i] = a + phic[i] - w;
            gn += grad[i] * grad[i];
            for (int64_t j = 0; j < K; j++) H[i * K + j] = invsigma[i * K + j];
            H[i * K + i] += w;
        }
        chol(K, H);
        chol_solve(K, H, grad);
        for (int64_t i = 0; i < K; i++) lambda_d[i] += grad[i];
        if (sqrt(gn) < ntol) break;
    }
}

#ifndef ORC_CTM_HELPERS_ONLY
typedef struct {
    double *phi, *phic, *grad, *H;
} ctm_ws;

static ctm_ws ws_alloc(int64_t K, int64_t mx)
{
    ctm_ws w;
    w.phi = (double *)malloc(sizeof(double) * K * mx);
    w.phic = (double *)malloc(sizeof(double) * K);
    w.grad = (double *)malloc(sizeof(double) * K);
    w.H = (double *)malloc(sizeof(double) * K * K);
    return w;
}
static void ws_free(ctm_ws w) { free(w.phi); free(w.phic); free(w.grad); free(w.H); }

/* CTM.jl:194-203  per-document inner loop; on exit ws.phi is the last phi (from lambda_old_d) */
static int ctm_doc_estep(int64_t K, int64_t Nd, const int64_t *terms, const int64_t *counts, double Cd, const double *beta,
                         const double *mu, const double *invsigma, double *lambda_d, double *lambda_old_d, double *vsq_d,
                         double *logzeta_d, int niter, double ntol, int viter, double vtol, ctm_ws w)
{
    int v;
    for (v = 0; v < viter; v++) {
        ctm_update_phi(K, Nd, terms, beta, lambda_d, w.phi);
        *logzeta_d = ctm_logzeta(K, lambda_d, vsq_d);
        ctm_update_vsq(K, Cd, invsigma, lambda_d, vsq_d, *logzeta_d, niter, ntol);
        ctm_update_lambda(K, Nd, counts, Cd, mu, invsigma, w.phi, lambda_d, lambda_old_d, vsq_d, *logzeta_d, niter, ntol, w.phic,
                          w.grad, w.H);
        double nrm = 0.0;
        for (int64_t i = 0; i < K; i++) { double df = lambda_d[i] - lambda_old_d[i]; nrm += df * df; }
        if (sqrt(nrm) < vtol) { v++; break; }
    }
    return v;
}

static int64_t max_len(int64_t M, const int64_t *off)
{
    int64_t mx = 1;
    for (int64_t d = 0; d < M; d++) if (off[d + 1] - off[d] > mx) mx = off[d + 1] - off[d];
    return mx;
}

/* CTM.jl:56-98  update_elbo!: lagged phi from beta_old / lambda_old; everything else current */
double orc_ctm_elbo(int64_t K, int64_t M, int64_t V, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts,
                    const double *mu, const double *invsigma, const double *beta, const double *beta_old, const double *lambda,
                    const double *lambda_old, const double *vsq, const double *logzeta, int nthreads)
{
    (void)V;
    double logdet_inv = 0.0;
    orc_spd_inv_logdet(K, invsigma, NULL, &logdet_inv);
    int64_t mx = max_len(M, N_cumsum);
    double elbo = 0.0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads) reduction(+ : elbo)
    {
        double *phi = (double *)malloc(sizeof(double) * K * mx);
        double *df = (double *)malloc(sizeof(double) * K);
#pragma omp for schedule(dynamic, 32)
        for (int64_t d = 0; d < M; d++) {
            int64_t o = N_cumsum[d], Nd = N_cumsum[d + 1] - o;
            const double *lam = lambda + K * d, *v = vsq + K * d;
            double Cd = 0.0;
            for (int64_t n = 0; n < Nd; n++) Cd += (double)counts[o + n];
            ctm_update_phi(K, Nd, terms + o, beta_old, lambda_old + K * d, phi);
            /* Elogpeta, CTM.jl:56-59 */
            double q = 0.0, dv = 0.0;
            for (int64_t i = 0; i < K; i++) { df[i] = lam[i] - mu[i]; dv += invsigma[i * K + i] * v[i]; }
            for (int64_t i = 0; i < K; i++) {
                double a = 0.0;
                for (int64_t j = 0; j < K; j++) a += invsigma[i * K + j] * df[j];
                q += df[i] * a;
            }
            double x = 0.5 * (logdet_inv - (double)K * log(2.0 * M_PI) - dv - q);
            /* Elogpz, CTM.jl:62-66 */
            double se = 0.0;
            for (int64_t i = 0; i < K; i++) se += exp(lam[i] + 0.5 * v[i] - logzeta[d]);
            double pz = 0.0, pw = 0.0, ez = 0.0;
            for (int64_t n = 0; n < Nd; n++) {
                const double *p = phi + K * n;
                const double *b = beta + K * terms[o + n];
                double c = (double)counts[o + n], a = 0.0, w = 0.0, e = 0.0;
                for (int64_t i = 0; i < K; i++) {
                    a += p[i] * lam[i];
                    w += p[i] * log(b[i] + ORC_EPS);            /* Elogpw, CTM.jl:69-73 */
                    if (p[i] > 0.0) e -= p[i] * log(p[i]);      /* -Elogqz, CTM.jl:83-87 */
                }
                pz += c * a; pw += c * w; ez += c * e;
            }
            x += pz - Cd * (se + logzeta[d] - 1.0) + pw + ez;
            /* -Elogqeta = entropy(MvNormal(lambda, diagm(vsq))), CTM.jl:76-80 */
            double ld = 0.0;
            for (int64_t i = 0; i < K; i++) ld += log(v[i]);
            x += 0.5 * ((double)K * (log(2.0 * M_PI) + 1.0) + ld);
            elbo += x;
        }
        free(phi);
        free(df);
    }
    return elbo;
}

/*
 * CTM.jl:185-217  train!.  In/out: mu[K], sigma[K*K], invsigma[K*K], beta[K*V], lambda[K*M], vsq[K*M], logzeta[M].
 * Out: beta_old[K*V], lambda_old[K*M], elbo_trace[iter+1] (NaN where not evaluated), sweeps_trace[iter].
 */
int orc_ctm_train(int64_t K, int64_t M, int64_t V, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts, double *mu,
                  double *sigma, double *invsigma, double *beta, double *beta_old, double *lambda, double *lambda_old, double *vsq,
                  double *logzeta, int iter, double tol, int niter, double ntol, int viter, double vtol, int checkelbo,
                  double *elbo_trace, int64_t *sweeps_trace, int *iters_done, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    int64_t KV = K * V, mx = max_len(M, N_cumsum);
    int all_empty = 1;
    for (int64_t d = 0; d < M; d++) if (N_cumsum[d + 1] > N_cumsum[d]) { all_empty = 0; break; }
    if (all_empty) iter = 0;
    memcpy(beta_old, beta, sizeof(double) * KV);
    memcpy(lambda_old, lambda, sizeof(double) * K * M);
    for (int k = 0; k <= iter; k++) elbo_trace[k] = NAN;
    int check = checkelbo > 0;
    double elbo = 0.0;
    if (check && checkelbo <= iter) {
        elbo = orc_ctm_elbo(K, M, V, N_cumsum, terms, counts, mu, invsigma, beta, beta_old, lambda, lambda_old, vsq, logzeta, nthreads);
        elbo_trace[0] = elbo;
    }
    double *beta_temp = (double *)calloc((size_t)KV * (size_t)nthreads, sizeof(double));
    double *S = (double *)malloc(sizeof(double) * K * K);
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
            ctm_ws w = ws_alloc(K, mx);
#pragma omp for schedule(dynamic, 32)
            for (int64_t d = 0; d < M; d++) {
                int64_t o = N_cumsum[d], Nd = N_cumsum[d + 1] - o;
                double Cd = 0.0;
                for (int64_t n = 0; n < Nd; n++) Cd += (double)counts[o + n];
                sweeps += ctm_doc_estep(K, Nd, terms + o, counts + o, Cd, beta, mu, invsigma, lambda + K * d, lambda_old + K * d,
                                        vsq + K * d, logzeta + d, niter, ntol, viter, vtol, w);
                /* CTM.jl:122-125 update_beta!(model, d) */
                for (int64_t n = 0; n < Nd; n++) {
                    double *b = bt + K * terms[o + n];
                    const double *p = w.phi + K * n;
                    double c = (double)counts[o + n];
                    for (int64_t i = 0; i < K; i++) b[i] += p[i] * c;
                }
            }
            ws_free(w);
        }
        for (int t = 1; t < nthreads; t++) {
            double *bt = beta_temp + (size_t)KV * t;
            for (int64_t q = 0; q < KV; q++) { beta_temp[q] += bt[q]; bt[q] = 0.0; }
        }
        if (sweeps_trace) sweeps_trace[k - 1] = sweeps;
        /* CTM.jl:114-118 update_beta!(model) */
        memcpy(beta_old, beta, sizeof(double) * KV);
        for (int64_t i = 0; i < K; i++) {
            double rs = 0.0;
            for (int64_t j = 0; j < V; j++) rs += beta_temp[K * j + i];
            for (int64_t j = 0; j < V; j++) beta[K * j + i] = beta_temp[K * j + i] / rs;
        }
        memset(beta_temp, 0, sizeof(double) * KV);
        /* CTM.jl:108-111 update_sigma! (uses the OLD mu: update_mu! runs after it, CTM.jl:207-208) */
        for (int64_t q = 0; q < K * K; q++) S[q] = 0.0;
        for (int64_t d = 0; d < M; d++) {
            const double *lam = lambda + K * d;
            for (int64_t i = 0; i < K; i++) {
                double di = lam[i] - mu[i];
                for (int64_t j = 0; j < K; j++) S[i * K + j] += di * (lam[j] - mu[j]);
                S[i * K + i] += vsq[K * d + i];
            }
        }
        for (int64_t q = 0; q < K * K; q++) sigma[q] = S[q] / (double)M;
        orc_spd_inv_logdet(K, sigma, invsigma, NULL);
        /* CTM.jl:102-104 update_mu! */
        for (int64_t i = 0; i < K; i++) {
            double s = 0.0;
            for (int64_t d = 0; d < M; d++) s += lambda[K * d + i];
            mu[i] = s / (double)M;
        }
        k_done = k;
        if (check && (k % checkelbo == 0)) {
            double e2 = orc_ctm_elbo(K, M, V, N_cumsum, terms, counts, mu, invsigma, beta, beta_old, lambda, lambda_old, vsq, logzeta,
                                     nthreads);
            double delta = e2 - elbo;
            elbo = e2;
            elbo_trace[k] = e2;
            if (delta < tol) break;
        }
    }
    free(beta_temp);
    free(S);
    if (iters_done) *iters_done = k_done;
    return 0;
}
#endif /* ORC_CTM_HELPERS_ONLY */
