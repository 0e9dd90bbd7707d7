/*
 * oracle/fctm_oracle.c -- fp64 CPU restatement of the reference's filtered CTM (src/fCTM.jl).
 *
 * TEST INFRASTRUCTURE ONLY (see lda_oracle.c); PARITY UNPINNED like the other oracles, cross-checked against the literal NumPy
 * transcription FCTMTwin (oracle/numpy_twin.py) to <= 1e-10 relative.  The Newton / Cholesky helpers are those of ctm_oracle.c
 * (update_lambda!, update_vsq!, update_logzeta! of fCTM.jl:181-226 are CTM.jl:129-171 verbatim); layout as there, tau / tau_old
 * flat over the CSR tokens.
 */
#define ORC_CTM_HELPERS_ONLY
#include "ctm_oracle.c"

static int64_t max_len(int64_t M, const int64_t *off)
{
    int64_t mx = 1;
    for (int64_t d = 0; d < M; d++) if (off[d + 1] - off[d] > mx) mx = off[d + 1] - off[d];
    return mx;
}

/* fCTM.jl:239-242  update_phi!: additive_logistic(tau' .* log.(@boink beta[:,terms]) .+ lambda[d], dims=1) */
static void fctm_update_phi(int64_t K, int64_t Nd, const int64_t *terms, const double *beta, const double *tau_d, const double *lambda_d,
                            double *phi)
{
    for (int64_t n = 0; n < Nd; n++) {
        const double *b = beta + K * terms[n];
        double *p = phi + K * n;
        double mx = -INFINITY, s = 0.0;
        for (int64_t i = 0; i < K; i++) { p[i] = tau_d[n] * log(b[i] + ORC_EPS) + lambda_d[i]; if (p[i] > mx) mx = p[i]; }
        for (int64_t i = 0; i < K; i++) { p[i] = exp(p[i] - mx); s += p[i]; }
        for (int64_t i = 0; i < K; i++) p[i] /= s;
    }
}

/* fCTM.jl:230-235  update_tau! */
static void fctm_update_tau(int64_t K, int64_t Nd, const int64_t *terms, const double *beta, const double *kappa, double eta,
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

static double fctm_bernoulli_entropy(double p)
{
    double p0 = 1.0 - p;
    if (p0 == 0.0 || p0 == 1.0) return 0.0;
    return -(p0 * log(p0) + p * log(p));
}

/* fCTM.jl:67-130  update_elbo! */
double orc_fctm_elbo(int64_t K, int64_t M, int64_t V, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts, double eta,
                     const double *mu, const double *invsigma, const double *kappa, const double *beta, const double *beta_old,
                     const double *lambda, const double *lambda_old, const double *vsq, const double *logzeta, const double *tau,
                     const double *tau_old, int nthreads)
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
            double Cd = 0.0, tc = 0.0;
            for (int64_t n = 0; n < Nd; n++) { Cd += (double)counts[o + n]; tc += tau[o + n] * (double)counts[o + n]; }
            fctm_update_phi(K, Nd, terms + o, beta_old, tau_old + o, lambda_old + K * d, phi);   /* fCTM.jl:125 */
            double q = 0.0, dv = 0.0; /* Elogpeta, fCTM.jl:67-70 */
            for (int64_t i = 0; i < K; i++) { df[i] = lam[i] - mu[i]; dv += invsigma[i * K + i] * v[i]; }
            for (int64_t i = 0; i < K; i++) {
                double a = 0.0;
                for (int64_t j = 0; j < K; j++) a += invsigma[i * K + j] * df[j];
                q += df[i] * a;
            }
            double x = 0.5 * (logdet_inv - (double)K * log(2.0 * M_PI) - dv - q);
            x += log(pow(eta, tc) * pow(1.0 - eta, Cd - tc) + ORC_EPS); /* Elogpc, fCTM.jl:73-77 */
            double se = 0.0;
            for (int64_t i = 0; i < K; i++) se += exp(lam[i] + 0.5 * v[i] - logzeta[d]);
            double pz = 0.0, pw = 0.0, ez = 0.0, ec = 0.0;
            for (int64_t n = 0; n < Nd; n++) {
                const double *p = phi + K * n;
                const double *b = beta + K * terms[o + n];
                double c = (double)counts[o + n], t = tau[o + n], a = 0.0, w = 0.0, e = 0.0;
                for (int64_t i = 0; i < K; i++) {
                    a += p[i] * lam[i];                          /* Elogpz, fCTM.jl:80-84 */
                    w += p[i] * log(b[i] + ORC_EPS);             /* Elogpw, fCTM.jl:87-91 */
                    if (p[i] > 0.0) e -= p[i] * log(p[i]);       /* -Elogqz, fCTM.jl:108-112 */
                }
                pz += c * a;
                pw += c * t * w + c * (1.0 - t) * log(kappa[terms[o + n]] + ORC_EPS);
                ez += c * e;
                ec += c * fctm_bernoulli_entropy(t);             /* -Elogqc, fCTM.jl:101-105 */
            }
            x += pz - Cd * (se + logzeta[d] - 1.0) + pw + ez + ec;
            double ld = 0.0; /* -Elogqeta, fCTM.jl:94-98 */
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
 * fCTM.jl:249-290  train!.  In/out: mu[K], sigma[K*K], invsigma[K*K], kappa[V], beta[K*V], lambda[K*M], vsq[K*M], logzeta[M], tau[nnz];
 * eta is constant (update_eta! is commented out in the reference's loop, fCTM.jl:279).
 * Out: kappa_old, beta_old, lambda_old, tau_old, elbo_trace[iter+1], sweeps_trace[iter], *iters_done.
 */
int orc_fctm_train(int64_t K, int64_t M, int64_t V, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts, double eta,
                   double *mu, double *sigma, double *invsigma, double *kappa, double *kappa_old, double *beta, double *beta_old,
                   double *lambda, double *lambda_old, double *vsq, double *logzeta, double *tau, double *tau_old, int iter, double tol,
                   int niter, double ntol, int viter, double vtol, int checkelbo, double *elbo_trace, int64_t *sweeps_trace,
                   int *iters_done, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    const int64_t KV = K * V, nnz = N_cumsum[M], mx = max_len(M, N_cumsum);
    int all_empty = 1;
    for (int64_t d = 0; d < M; d++) if (N_cumsum[d + 1] > N_cumsum[d]) { all_empty = 0; break; }
    if (all_empty) iter = 0;
    memcpy(beta_old, beta, sizeof(double) * KV);
    memcpy(kappa_old, kappa, sizeof(double) * V);
    memcpy(lambda_old, lambda, sizeof(double) * K * M);
    memcpy(tau_old, tau, sizeof(double) * nnz);
    for (int k = 0; k <= iter; k++) elbo_trace[k] = NAN;
    int check = checkelbo > 0;
    double elbo = 0.0;
    if (check && checkelbo <= iter) {
        elbo = orc_fctm_elbo(K, M, V, N_cumsum, terms, counts, eta, mu, invsigma, kappa, beta, beta_old, lambda, lambda_old, vsq, logzeta,
                             tau, tau_old, nthreads);
        elbo_trace[0] = elbo;
    }
    double *temp = (double *)calloc((size_t)(KV + V) * (size_t)nthreads, sizeof(double));
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
            double *bt = temp + (size_t)(KV + V) * tid, *kt = bt + KV;
            double *phi = (double *)malloc(sizeof(double) * K * mx);
            double *phic = (double *)malloc(sizeof(double) * K), *grad = (double *)malloc(sizeof(double) * K);
            double *H = (double *)malloc(sizeof(double) * K * K);
#pragma omp for schedule(dynamic, 32)
            for (int64_t d = 0; d < M; d++) {
                int64_t o = N_cumsum[d], Nd = N_cumsum[d + 1] - o;
                double Cd = 0.0;
                for (int64_t n = 0; n < Nd; n++) Cd += (double)counts[o + n];
                double *lam = lambda + K * d, *lo = lambda_old + K * d, *v = vsq + K * d;
                int sw;
                for (sw = 0; sw < viter; sw++) { /* fCTM.jl:258-268 */
                    fctm_update_phi(K, Nd, terms + o, beta, tau + o, lam, phi);
                    fctm_update_tau(K, Nd, terms + o, beta, kappa, eta, phi, tau + o, tau_old + o);
                    logzeta[d] = ctm_logzeta(K, lam, v);
                    ctm_update_lambda(K, Nd, counts + o, Cd, mu, invsigma, phi, lam, lo, v, logzeta[d], niter, ntol, phic, grad, H);
                    ctm_update_vsq(K, Cd, invsigma, lam, v, logzeta[d], niter, ntol);
                    double nrm = 0.0;
                    for (int64_t i = 0; i < K; i++) { double df = lam[i] - lo[i]; nrm += df * df; }
                    if (sqrt(nrm) < vtol) { sw++; break; }
                }
                sweeps += sw;
                for (int64_t n = 0; n < Nd; n++) {
                    double *b = bt + K * terms[o + n];
                    const double *p = phi + K * n;
                    double c = (double)counts[o + n], t = tau[o + n];
                    for (int64_t i = 0; i < K; i++) b[i] += p[i] * (t * c); /* fCTM.jl:175-178 */
                    kt[terms[o + n]] += (1.0 - t) * c;                      /* fCTM.jl:162-165 */
                }
            }
            free(phi); free(phic); free(grad); free(H);
        }
        for (int t = 1; t < nthreads; t++) {
            double *bt = temp + (size_t)(KV + V) * t;
            for (int64_t q = 0; q < KV + V; q++) { temp[q] += bt[q]; bt[q] = 0.0; }
        }
        if (sweeps_trace) sweeps_trace[k - 1] = sweeps;
        memcpy(beta_old, beta, sizeof(double) * KV); /* fCTM.jl:167-171 */
        for (int64_t i = 0; i < K; i++) {
            double rs = 0.0;
            for (int64_t j = 0; j < V; j++) rs += temp[K * j + i];
            for (int64_t j = 0; j < V; j++) beta[K * j + i] = temp[K * j + i] / rs;
        }
        memcpy(kappa_old, kappa, sizeof(double) * V); /* fCTM.jl:154-158 */
        {
            double ks = 0.0;
            for (int64_t j = 0; j < V; j++) ks += temp[KV + j];
            for (int64_t j = 0; j < V; j++) kappa[j] = temp[KV + j] / ks;
        }
        memset(temp, 0, sizeof(double) * (KV + V));
        for (int64_t q = 0; q < K * K; q++) S[q] = 0.0; /* update_sigma! with the OLD mu (fCTM.jl:147-150, call order :276-277) */
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
        for (int64_t i = 0; i < K; i++) { /* update_mu!, fCTM.jl:141-143 */
            double s = 0.0;
            for (int64_t d = 0; d < M; d++) s += lambda[K * d + i];
            mu[i] = s / (double)M;
        }
        k_done = k;
        if (check && (k % checkelbo == 0)) {
            double e2 = orc_fctm_elbo(K, M, V, N_cumsum, terms, counts, eta, mu, invsigma, kappa, beta, beta_old, lambda, lambda_old, vsq,
                                      logzeta, tau, tau_old, nthreads);
            double delta = e2 - elbo;
            elbo = e2;
            elbo_trace[k] = e2;
            if (delta < tol) break;
        }
    }
    free(temp);
    free(S);
    if (iters_done) *iters_done = k_done;
    return 0;
}
