"""Independent NumPy/SciPy restatement of the reference's CPU models -- TEST INFRASTRUCTURE ONLY.

Second opinion for the C oracle (``*_oracle.c``): written straight from the Julia source with
scipy.special for psi / psi' / ln Gamma, per-document Python loops (small cases only).
The two restatements must agree to <= 1e-12 relative (tests/test_oracle_*.py).

Arrays are (V, K) / (M, K) C-order views of the reference's column-major K x V / K x M.
"""
from __future__ import annotations

import numpy as np
from scipy.special import digamma, gammaln, polygamma

EPSILON = 2.0 ** -99  # utils.jl:3  eps(1e-14)


def _finite(x):  # utils.jl:107
    return np.sign(x) * np.minimum(np.abs(x), np.finfo(np.float64).max)


class LDATwin:
    """src/LDA.jl, line for line."""

    def __init__(self, N_cumsum, terms, counts, K, V, beta, alpha=None):
        self.K, self.V = K, V
        self.M = len(N_cumsum) - 1
        self.off = np.asarray(N_cumsum, dtype=np.int64)
        self.terms = np.asarray(terms, dtype=np.int64)
        self.counts = np.asarray(counts, dtype=np.float64)
        self.alpha = np.ones(K) if alpha is None else np.array(alpha, dtype=np.float64)
        self.beta = np.array(beta, dtype=np.float64).reshape(V, K)
        self.beta_old = self.beta.copy()
        self.beta_temp = np.zeros((V, K))
        self.Elogtheta = np.full((self.M, K), -np.euler_gamma - digamma(K))  # LDA.jl:38
        self.Elogtheta_old = self.Elogtheta.copy()
        self.gamma = np.ones((self.M, K))
        self.phi = None
        self.elbo = 0.0

    def _doc(self, d):
        s = slice(self.off[d], self.off[d + 1])
        return self.terms[s], self.counts[s]

    # LDA.jl:150-154
    def update_phi(self, d):
        terms, _ = self._doc(d)
        phi = EPSILON + self.beta[terms] * np.exp(self.Elogtheta[d])[None, :]
        self.phi = phi / phi.sum(axis=1, keepdims=True)

    # LDA.jl:143-146
    def update_gamma(self, d):
        _, counts = self._doc(d)
        self.gamma[d] = EPSILON + (self.alpha + counts @ self.phi)

    # LDA.jl:136-139
    def update_Elogtheta(self, d):
        self.Elogtheta_old[d] = self.Elogtheta[d]
        self.Elogtheta[d] = digamma(self.gamma[d]) - digamma(self.gamma[d].sum())

    # LDA.jl:129-132
    def update_beta_doc(self, d):
        terms, counts = self._doc(d)
        self.beta_temp[terms] += self.phi * counts[:, None]

    # LDA.jl:121-125
    def update_beta(self):
        self.beta_old = self.beta
        self.beta = self.beta_temp / self.beta_temp.sum(axis=0, keepdims=True)
        self.beta_temp = np.zeros((self.V, self.K))

    # LDA.jl:97-118
    def update_alpha(self, niter, ntol):
        Elogtheta_sum = self.Elogtheta.sum(axis=0)
        K, M = self.K, self.M
        nu = float(K)
        for _ in range(niter):
            rho = 1.0
            a = self.alpha
            grad = nu / a + M * (digamma(a.sum()) - digamma(a)) + Elogtheta_sum
            h_inv = -1.0 / (M * polygamma(1, a) + nu / a**2)
            p = (grad - np.dot(grad, h_inv) / (1.0 / (M * polygamma(1, a.sum())) + h_inv.sum())) * h_inv
            while np.min(a - rho * p) < 0:
                rho *= 0.5
            self.alpha = np.sign(a) * np.minimum(np.abs(a - rho * p), np.finfo(np.float64).max)
            if (rho * np.linalg.norm(grad) < ntol) and (nu / K < ntol):
                break
            nu *= 0.5
        self.alpha = self.alpha + EPSILON

    # LDA.jl:50-93
    def update_elbo(self):
        elbo = 0.0
        a = self.alpha
        for d in range(self.M):
            terms, counts = self._doc(d)
            phi = EPSILON + self.beta_old[terms] * np.exp(self.Elogtheta_old[d])[None, :]
            phi = phi / phi.sum(axis=1, keepdims=True)
            Et, g = self.Elogtheta[d], self.gamma[d]
            Elogptheta = _finite(gammaln(a.sum())) - _finite(gammaln(a).sum()) + np.dot(a - 1, Et)
            Elogpz = np.dot(counts @ phi, Et)
            Elogpw = np.sum((phi * np.log(self.beta[terms] + EPSILON)) * counts[:, None])
            if self.K == 1:
                ent_dir = 0.0
            else:  # utils.jl:163-180
                ent_dir = (gammaln(g).sum() - gammaln(g.sum()) + (g.sum() - self.K) * digamma(g.sum())
                           - np.dot(g - 1.0, digamma(g)))
            with np.errstate(divide="ignore", invalid="ignore"):
                plogp = np.where(phi > 0, phi * np.log(phi), 0.0)
            ent_z = -(plogp.sum(axis=1) * counts).sum()
            elbo += Elogptheta + Elogpz + Elogpw + ent_dir + ent_z
        self.elbo = elbo
        return elbo

    # LDA.jl:161-191 (+ check_elbo!, modelutils.jl:574-585)
    def train(self, iter=150, tol=1.0, niter=1000, ntol=None, viter=10, vtol=None, checkelbo=1):
        ntol = 1.0 / self.K**2 if ntol is None else ntol
        vtol = 1.0 / self.K**2 if vtol is None else vtol
        trace = [np.nan] * (iter + 1)
        if checkelbo <= iter:
            trace[0] = self.update_elbo()
        for k in range(1, iter + 1):
            for d in range(self.M):
                for _ in range(viter):
                    self.update_phi(d)
                    self.update_gamma(d)
                    self.update_Elogtheta(d)
                    if np.linalg.norm(self.Elogtheta[d] - self.Elogtheta_old[d]) < vtol:
                        break
                self.update_beta_doc(d)
            self.update_beta()
            self.update_alpha(niter, ntol)
            if k % checkelbo == 0:
                old = self.elbo
                delta = self.update_elbo() - old
                trace[k] = self.elbo
                if delta < tol:
                    break
        return np.array(trace)


class CTMTwin:
    """src/CTM.jl, line for line (numpy.linalg for `\\`, inv, logdet)."""

    def __init__(self, N_cumsum, terms, counts, K, V, beta):
        self.K, self.V = K, V
        self.M = len(N_cumsum) - 1
        self.off = np.asarray(N_cumsum, dtype=np.int64)
        self.terms = np.asarray(terms, dtype=np.int64)
        self.counts = np.asarray(counts, dtype=np.float64)
        self.C = np.array([self.counts[self.off[d]:self.off[d + 1]].sum() for d in range(self.M)])
        self.mu = np.zeros(K)                                   # CTM.jl:38-49
        self.sigma = np.eye(K)
        self.invsigma = np.eye(K)
        self.beta = np.array(beta, dtype=np.float64).reshape(V, K)
        self.beta_old = self.beta.copy()
        self.beta_temp = np.zeros((V, K))
        self.lam = np.zeros((self.M, K))
        self.lam_old = np.zeros((self.M, K))
        self.vsq = np.ones((self.M, K))
        self.logzeta = np.full(self.M, 0.5)
        self.phi = None
        self.elbo = 0.0

    def _doc(self, d):
        s = slice(self.off[d], self.off[d + 1])
        return self.terms[s], self.counts[s]

    @staticmethod
    def _softmax_rows(x):
        x = np.exp(x - x.max(axis=1, keepdims=True))
        return x / x.sum(axis=1, keepdims=True)

    def update_phi(self, d):  # CTM.jl:175-178
        terms, _ = self._doc(d)
        with np.errstate(divide="ignore"):
            self.phi = self._softmax_rows(np.log(self.beta[terms]) + self.lam[d][None, :])

    def update_logzeta(self, d):  # CTM.jl:169-171
        x = self.lam[d] + 0.5 * self.vsq[d]
        self.logzeta[d] = x.max() + np.log(np.exp(x - x.max()).sum())

    def update_vsq(self, d, niter, ntol):  # CTM.jl:146-165
        for i in range(self.K):
            for _ in range(niter):
                rho = 1.0
                ex = self.C[d] * np.exp(self.lam[d, i] + 0.5 * self.vsq[d, i] - self.logzeta[d])
                grad = -0.5 * (self.invsigma[i, i] + ex - 1.0 / self.vsq[d, i])
                invhess = -1.0 / (0.25 * ex + 0.5 / self.vsq[d, i] ** 2)
                p = invhess * grad
                while self.vsq[d, i] - rho * p <= 0:
                    rho *= 0.5
                self.vsq[d, i] -= rho * p
                if rho * abs(grad) < ntol:
                    break
        self.vsq[d] += EPSILON

    def update_lambda(self, d, niter, ntol):  # CTM.jl:129-142
        self.lam_old[d] = self.lam[d]
        _, counts = self._doc(d)
        phic = counts @ self.phi
        for _ in range(niter):
            w = self.C[d] * np.exp(self.lam[d] + 0.5 * self.vsq[d] - self.logzeta[d])
            grad = self.invsigma @ (self.mu - self.lam[d]) + phic - w
            H = self.invsigma + np.diag(w)
            self.lam[d] = self.lam[d] + np.linalg.solve(H, grad)
            if np.linalg.norm(grad) < ntol:
                break

    def update_elbo(self):  # CTM.jl:56-98
        elbo = 0.0
        K = self.K
        _, logdet = np.linalg.slogdet(self.invsigma)
        for d in range(self.M):
            terms, counts = self._doc(d)
            with np.errstate(divide="ignore"):
                phi = self._softmax_rows(np.log(self.beta_old[terms]) + self.lam_old[d][None, :])
            lam, v = self.lam[d], self.vsq[d]
            df = lam - self.mu
            Elogpeta = 0.5 * (logdet - K * np.log(2 * np.pi) - np.dot(np.diag(self.invsigma), v) - df @ self.invsigma @ df)
            Elogpz = np.dot(phi @ lam, counts) - self.C[d] * (np.exp(lam + 0.5 * v - self.logzeta[d]).sum() + self.logzeta[d] - 1)
            Elogpw = np.sum((phi * np.log(self.beta[terms] + EPSILON)) * counts[:, None])
            ent_eta = 0.5 * (K * (np.log(2 * np.pi) + 1) + np.log(v).sum())
            with np.errstate(divide="ignore", invalid="ignore"):
                plogp = np.where(phi > 0, phi * np.log(phi), 0.0)
            ent_z = -(plogp.sum(axis=1) * counts).sum()
            elbo += Elogpeta + Elogpz + Elogpw + ent_eta + ent_z
        self.elbo = elbo
        return elbo

    def train(self, iter=150, tol=1.0, niter=1000, ntol=None, viter=10, vtol=None, checkelbo=1):  # CTM.jl:185-217
        K = self.K
        ntol = 1.0 / K**2 if ntol is None else ntol
        vtol = 1.0 / K**2 if vtol is None else vtol
        trace = [np.nan] * (iter + 1)
        if checkelbo <= iter:
            trace[0] = self.update_elbo()
        for k in range(1, iter + 1):
            for d in range(self.M):
                terms, counts = self._doc(d)
                for _ in range(viter):
                    self.update_phi(d)
                    self.update_logzeta(d)
                    self.update_vsq(d, niter, ntol)
                    self.update_lambda(d, niter, ntol)
                    if np.linalg.norm(self.lam[d] - self.lam_old[d]) < vtol:
                        break
                self.beta_temp[terms] += self.phi * counts[:, None]      # CTM.jl:122-125
            self.beta_old = self.beta                                    # CTM.jl:114-118
            self.beta = self.beta_temp / self.beta_temp.sum(axis=0, keepdims=True)
            self.beta_temp = np.zeros((self.V, K))
            dl = self.lam - self.mu[None, :]                             # CTM.jl:108-111 (old mu)
            self.sigma = (np.diag(self.vsq.sum(axis=0)) + dl.T @ dl) / self.M
            self.invsigma = np.linalg.inv(self.sigma)
            self.mu = self.lam.sum(axis=0) / self.M                      # CTM.jl:102-104
            if k % checkelbo == 0:
                old = self.elbo
                delta = self.update_elbo() - old
                trace[k] = self.elbo
                if delta < tol:
                    break
        return np.array(trace)
