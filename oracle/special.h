/*
 * oracle/special.h -- fp64 special functions for the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * The reference takes digamma / trigamma / loggamma from SpecialFunctions.jl
 * (Project.toml:16, compat "0.8, 0.9, 0.10"; not vendored under /root/reference),
 * used at LDA.jl:38,51,103-105,138 and CTPF.jl:116-247.  They are the textbook
 * functions psi(x), psi'(x), ln Gamma(x); restated here from their published
 * recurrences + asymptotic (Bernoulli) series and pinned against scipy.special in
 * tests/test_oracle_special.py (<= 4e-15 relative on (1e-3, 1e4)).
 */
#ifndef TMVB_ORACLE_SPECIAL_H
#define TMVB_ORACLE_SPECIAL_H

#include <math.h>

/* psi(x), x > 0: shift up to x >= 10 with psi(x) = psi(x+1) - 1/x, then
 * psi(x) ~ ln x - 1/(2x) - sum_k B_2k / (2k x^2k). */
static inline double orc_digamma(double x)
{
    double r = 0.0;
    while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
    double t = 1.0 / x, t2 = t * t;
    double s = t2 * (1.0 / 12 - t2 * (1.0 / 120 - t2 * (1.0 / 252 - t2 * (1.0 / 240 -
               t2 * (1.0 / 132 - t2 * (691.0 / 32760 - t2 * (1.0 / 12)))))));
    return r + log(x) - 0.5 * t - s;
}

/* psi'(x), x > 0: psi'(x) = psi'(x+1) + 1/x^2, then
 * psi'(x) ~ 1/x + 1/(2x^2) + sum_k B_2k / x^(2k+1). */
static inline double orc_trigamma(double x)
{
    double r = 0.0;
    while (x < 10.0) { r += 1.0 / (x * x); x += 1.0; }
    double t = 1.0 / x, t2 = t * t;
    double s = t * (1.0 + 0.5 * t + t2 * (1.0 / 6 - t2 * (1.0 / 30 - t2 * (1.0 / 42 -
               t2 * (1.0 / 30 - t2 * (5.0 / 66 - t2 * (691.0 / 2730 - t2 * (7.0 / 6))))))));
    return r + s;
}

/* ln Gamma(x), x > 0 (thread-safe libm entry point). */
static inline double orc_lgamma(double x)
{
    int sg;
    return lgamma_r(x, &sg);
}

#endif
