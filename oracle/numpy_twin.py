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

    def update_elbo_device_form(self):
        """The same ELBO through the decomposition the CUDA path evaluates: valid right after an E-step +
        M-step + alpha update, when Elogtheta = psi(gamma) - psi(sum gamma) and gamma = alpha_estep + phi*c + eps hold.
          per document   sum_i lnG(gamma_i) - lnG(sum gamma) + sum_n c_n ln s_n - sum_i (gamma_i - alpha_estep_i) Elogtheta_old_i
          K-vectors      M (lnG(sum alpha) - sum lnG(alpha)) + sum_i (alpha_i - alpha_estep_i - eps) sum_d Elogtheta_di
          K x V          sum_ij S_ij [ln(beta_ij + eps) - ln(beta_old_ij + eps)],  S = the statistics of the last E-step
        No logarithm per (token, topic); agreement with update_elbo() pins the identity on the CPU."""
        a, ae = self.alpha, self.alpha_estep
        S = np.zeros((self.V, self.K))
        docs = 0.0
        for d in range(self.M):
            terms, counts = self._doc(d)
            u = EPSILON + self.beta_old[terms] * np.exp(self.Elogtheta_old[d])[None, :]
            s = u.sum(axis=1)
            S[terms] += (u / s[:, None]) * counts[:, None]
            g = self.gamma[d]
            docs += gammaln(g).sum() - gammaln(g.sum()) + np.dot(counts, np.log(s)) - np.dot(g - ae, self.Elogtheta_old[d])
        lin = np.dot(a - ae - EPSILON, self.Elogtheta.sum(axis=0))
        elbo_w = np.sum(S * (np.log(self.beta + EPSILON) - np.log(self.beta_old + EPSILON)))
        return docs + self.M * (gammaln(a.sum()) - gammaln(a).sum()) + lin + elbo_w

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
            self.alpha_estep = self.alpha.copy()   # not part of LDA.jl: kept for update_elbo_device_form
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

    def update_elbo_device_form(self):
        """CTM.jl:56-98 without a logarithm per (token, topic) -- the decomposition a device kernel can evaluate from what the
        E-step already holds (cf. LDATwin.update_elbo_device_form).  With u_ni = beta_old_i,w exp(lambda_old_i - mx_d),
        s_n = sum_i u_ni, phi = u / s:
          -Elogqz  = sum_n c_n ln s_n - sum_i (phi c)_i (lambda_old_i - mx_d) - sum_ij S_ij ln beta_old_ij
          Elogpw   = sum_ij S_ij ln(beta_ij + eps)                       (both K x V sums over the statistics S)
          Elogpz   = (phi c) . lambda - C_d (sum_i exp(lambda_i + v_i / 2 - ln zeta) + ln zeta - 1)
        Elogpeta and the Gaussian entropy are K-vector algebra per document."""
        K = self.K
        _, logdet = np.linalg.slogdet(self.invsigma)
        S = np.zeros((self.V, K))
        docs = 0.0
        for d in range(self.M):
            terms, counts = self._doc(d)
            mx = self.lam_old[d].max()
            u = self.beta_old[terms] * np.exp(self.lam_old[d] - mx)[None, :]
            s = u.sum(axis=1)
            phic = counts @ (u / s[:, None])
            S[terms] += (u / s[:, None]) * counts[:, None]
            lam, v = self.lam[d], self.vsq[d]
            df = lam - self.mu
            docs += 0.5 * (logdet - K * np.log(2 * np.pi) - np.dot(np.diag(self.invsigma), v) - df @ self.invsigma @ df)
            docs += np.dot(phic, lam) - self.C[d] * (np.exp(lam + 0.5 * v - self.logzeta[d]).sum() + self.logzeta[d] - 1)
            docs += 0.5 * (K * (np.log(2 * np.pi) + 1) + np.log(v).sum())
            docs += np.dot(counts, np.log(s)) - np.dot(phic, self.lam_old[d] - mx)
        pos = S > 0
        with np.errstate(divide="ignore"):
            glob = np.sum(S[pos] * (np.log(self.beta[pos] + EPSILON) - np.log(self.beta_old[pos])))
        return docs + glob

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


class CTPFTwin:
    """src/CTPF.jl restated with NumPy.  The ELBO uses the closed form the Binomial * lnGamma sums of
    Elogpya/Elogpyb/Elogpz (CTPF.jl:111-141) and of Distributions' entropy(Multinomial) (CTPF.jl:179-194) collapse
    to -- per token -lnG(c+1) + c H(phi_n), per reader -lnG(r+1) + r H(xi_r) -- so agreement with the C oracle
    (which evaluates the long form) also checks that identity."""

    def __init__(self, N_cumsum, terms, counts, R_cumsum, readers, ratings, K, V, U, alef, hyp=None):
        self.K, self.V, self.U = K, V, U
        self.M = len(N_cumsum) - 1
        self.off, self.roff = np.asarray(N_cumsum, np.int64), np.asarray(R_cumsum, np.int64)
        self.terms, self.counts = np.asarray(terms, np.int64), np.asarray(counts, np.float64)
        self.readers, self.ratings = np.asarray(readers, np.int64), np.asarray(ratings, np.float64)
        self.a, self.b, self.c, self.d, self.e, self.f, self.g, self.h = (0.1,) * 8 if hyp is None else hyp
        self.alef = np.array(alef, dtype=np.float64).reshape(V, K)          # CTPF.jl:83-100
        self.he = np.ones((U, K))
        self.bet, self.vav, self.dalet, self.het = np.ones(K), np.ones(K), np.ones(K), np.ones(K)
        self.gimel, self.zayin = np.ones((self.M, K)), np.ones((self.M, K))
        for n in ("alef", "he", "bet", "vav", "dalet", "het", "gimel", "zayin"):
            setattr(self, n + "_old", getattr(self, n).copy())
        self.alef_temp, self.he_temp = np.full((V, K), self.a), np.full((U, K), self.e)
        self.elbo = 0.0

    @staticmethod
    def _softmax_rows(x):
        x = np.exp(x - x.max(axis=1, keepdims=True))
        return x / x.sum(axis=1, keepdims=True)

    def _phi(self, d, alef, gimel, dalet, bet):  # CTPF.jl:327-330
        t = self.terms[self.off[d]:self.off[d + 1]]
        return self._softmax_rows((digamma(gimel[d]) - np.log(dalet) - np.log(bet))[None, :] + digamma(alef[t]))

    def _xi(self, d, he, gimel, zayin, dalet, het, vav):  # CTPF.jl:334-337
        r = self.readers[self.roff[d]:self.roff[d + 1]]
        ph = digamma(he[r])
        xa = (digamma(gimel[d]) - np.log(dalet) - np.log(vav))[None, :] + ph
        xb = (digamma(zayin[d]) - np.log(het) - np.log(vav))[None, :] + ph
        return self._softmax_rows(np.concatenate([xa, xb], axis=1)) if len(r) else np.zeros((0, 2 * self.K))

    @staticmethod
    def _ent_gamma(alpha, theta):
        return alpha + np.log(theta) + gammaln(alpha) + (1 - alpha) * digamma(alpha)

    def update_elbo(self):  # CTPF.jl:111-247
        K = self.K
        a, b, c, dd, e, f, g, h = self.a, self.b, self.c, self.d, self.e, self.f, self.g, self.h
        x = self.V * K * (a * np.log(b) - gammaln(a)) + np.sum((a - 1) * (digamma(self.alef) - np.log(self.bet)[None, :]) - b * self.alef / self.bet[None, :])
        x += np.sum(self._ent_gamma(self.alef, 1.0 / self.bet[None, :]))
        x += self.U * K * (e * np.log(f) - gammaln(e)) + np.sum((e - 1) * (digamma(self.he) - np.log(self.vav)[None, :]) - f * self.he / self.vav[None, :])
        x += np.sum(self._ent_gamma(self.he, 1.0 / self.vav[None, :]))
        alsum, hesum = self.alef.sum(axis=0), self.he.sum(axis=0)
        for d in range(self.M):
            t, cn = self.terms[self.off[d]:self.off[d + 1]], self.counts[self.off[d]:self.off[d + 1]]
            r, ra = self.readers[self.roff[d]:self.roff[d + 1]], self.ratings[self.roff[d]:self.roff[d + 1]]
            phi = self._phi(d, self.alef_old, self.gimel_old, self.dalet_old, self.bet_old)
            xi = self._xi(d, self.he_old, self.gimel_old, self.zayin_old, self.dalet_old, self.het_old, self.vav_old)
            gm, zy = self.gimel[d], self.zayin[d]
            x -= np.dot(gm / (self.dalet * self.vav), hesum) + np.dot(zy / (self.het * self.vav), hesum) + np.dot(gm / (self.dalet * self.bet), alsum)
            ph = digamma(self.he[r])
            x += np.sum(ra[:, None] * xi[:, :K] * ((digamma(gm) - np.log(self.dalet) - np.log(self.vav))[None, :] + ph))
            x += np.sum(ra[:, None] * xi[:, K:] * ((digamma(zy) - np.log(self.het) - np.log(self.vav))[None, :] + ph))
            x += np.sum(cn[:, None] * phi * ((digamma(gm) - np.log(self.dalet) - np.log(self.bet))[None, :] + digamma(self.alef[t])))
            with np.errstate(divide="ignore", invalid="ignore"):
                hx = -np.where(xi > 0, xi * np.log(xi), 0.0).sum(axis=1)
                hp = -np.where(phi > 0, phi * np.log(phi), 0.0).sum(axis=1)
            x += np.sum(ra * hx - gammaln(ra + 1)) + np.sum(cn * hp - gammaln(cn + 1))
            x += K * (c * np.log(dd) - gammaln(c)) + np.sum((c - 1) * (digamma(gm) - np.log(self.dalet)) - dd * gm / self.dalet)
            x += K * (g * np.log(h) - gammaln(g)) + np.sum((g - 1) * (digamma(zy) - np.log(self.het)) - h * zy / self.het)
            x += np.sum(self._ent_gamma(gm, 1.0 / self.dalet)) + np.sum(self._ent_gamma(zy, 1.0 / self.het))
        self.elbo = x
        return x

    def update_elbo_device_form(self):
        """update_elbo() with the two entropy sums written without a logarithm per (token, topic) / (reader, topic) -- valid
        after an E-step + M-step, when alef - a and he - e ARE the statistics of the last phi / xi.  With
        q = psi(gimel_old) - ln dalet_old - ln bet_old and lse_n = ln sum_i exp(q_i + psi(alef_old[w_n, i])):
          sum_n c_n H(phi_n) = sum_n c_n lse_n - (phi c) . q - sum_wi (alef - a)_wi psi(alef_old)_wi
        and for the 2K-way xi with qa, qb (CTPF.jl:334-337) and the K x U table psi(he_old):
          sum_r r H(xi_r) = sum_r r lse_r - (xi_a r) . qa - (xi_b r) . qb - sum_ui (he - e)_ui psi(he_old)_ui ."""
        K = self.K
        full = self.update_elbo()
        ent_lit, ent_dev = 0.0, 0.0
        for d in range(self.M):
            t, cn = self.terms[self.off[d]:self.off[d + 1]], self.counts[self.off[d]:self.off[d + 1]]
            r, ra = self.readers[self.roff[d]:self.roff[d + 1]], self.ratings[self.roff[d]:self.roff[d + 1]]
            phi = self._phi(d, self.alef_old, self.gimel_old, self.dalet_old, self.bet_old)
            xi = self._xi(d, self.he_old, self.gimel_old, self.zayin_old, self.dalet_old, self.het_old, self.vav_old)
            with np.errstate(divide="ignore", invalid="ignore"):
                ent_lit += np.sum(ra * -np.where(xi > 0, xi * np.log(xi), 0.0).sum(axis=1)) + np.sum(cn * -np.where(phi > 0, phi * np.log(phi), 0.0).sum(axis=1))
            q = digamma(self.gimel_old[d]) - np.log(self.dalet_old) - np.log(self.bet_old)
            x = q[None, :] + digamma(self.alef_old[t])
            lse = x.max(axis=1) + np.log(np.exp(x - x.max(axis=1, keepdims=True)).sum(axis=1))
            ent_dev += np.dot(cn, lse) - np.dot(cn @ phi, q)
            if len(r):
                qa = digamma(self.gimel_old[d]) - np.log(self.dalet_old) - np.log(self.vav_old)
                qb = digamma(self.zayin_old[d]) - np.log(self.het_old) - np.log(self.vav_old)
                ph = digamma(self.he_old[r])
                y = np.concatenate([qa[None, :] + ph, qb[None, :] + ph], axis=1)
                lse_r = y.max(axis=1) + np.log(np.exp(y - y.max(axis=1, keepdims=True)).sum(axis=1))
                ent_dev += np.dot(ra, lse_r) - np.dot(ra @ xi[:, :K], qa) - np.dot(ra @ xi[:, K:], qb)
        ent_dev -= np.sum((self.alef - self.a) * digamma(self.alef_old)) + np.sum((self.he - self.e) * digamma(self.he_old))
        return full - ent_lit + ent_dev

    def train(self, iter=150, tol=1.0, viter=10, vtol=None, checkelbo=1):  # CTPF.jl:344-371
        K = self.K
        vtol = 1.0 / K**2 if vtol is None else vtol
        trace = [np.nan] * (iter + 1)
        if checkelbo <= iter:
            trace[0] = self.update_elbo()
        for k in range(1, iter + 1):
            for d in range(self.M):
                t, cn = self.terms[self.off[d]:self.off[d + 1]], self.counts[self.off[d]:self.off[d + 1]]
                r, ra = self.readers[self.roff[d]:self.roff[d + 1]], self.ratings[self.roff[d]:self.roff[d + 1]]
                for _ in range(viter):
                    xi = self._xi(d, self.he, self.gimel, self.zayin, self.dalet, self.het, self.vav)
                    phi = self._phi(d, self.alef, self.gimel, self.dalet, self.bet)
                    self.zayin_old[d] = self.zayin[d]
                    self.zayin[d] = self.g + ra @ xi[:, K:]
                    self.gimel_old[d] = self.gimel[d]
                    self.gimel[d] = self.c + cn @ phi + ra @ xi[:, :K]
                    if np.linalg.norm(self.gimel[d] - self.gimel_old[d]) < vtol:
                        break
                np.add.at(self.he_temp, r, (xi[:, :K] + xi[:, K:]) * ra[:, None])
                self.alef_temp[t] += phi * cn[:, None]
            self.he_old, self.he, self.he_temp = self.he, self.he_temp, np.full((self.U, K), self.e)
            self.alef_old, self.alef, self.alef_temp = self.alef, self.alef_temp, np.full((self.V, K), self.a)
            self.dalet_old, self.dalet = self.dalet, self.d + self.alef.sum(axis=0) / self.bet + self.he.sum(axis=0) / self.vav
            self.het_old, self.het = self.het, self.h + self.he.sum(axis=0) / self.vav
            self.bet_old, self.bet = self.bet, self.b + self.gimel.sum(axis=0) / self.dalet
            self.vav_old, self.vav = self.vav, self.f + self.gimel.sum(axis=0) / self.dalet + self.zayin.sum(axis=0) / self.het
            if k % checkelbo == 0:
                old = self.elbo
                delta = self.update_elbo() - old
                trace[k] = self.elbo
                if delta < tol:
                    break
        return np.array(trace)


class FLDATwin(LDATwin):
    """src/fLDA.jl, line for line (filtered LDA: per-token Bernoulli tau, corpus distribution kappa, mixing weight eta)."""

    def __init__(self, N_cumsum, terms, counts, K, V, beta, kappa, alpha=None, eta=0.5):
        super().__init__(N_cumsum, terms, counts, K, V, beta, alpha)
        self.eta = float(eta)                                              # fLDA.jl:39
        self.kappa = np.array(kappa, dtype=np.float64)
        self.kappa_old = self.kappa.copy()
        self.kappa_temp = np.zeros(V)
        self.tau = np.full(len(self.terms), self.eta)                      # fLDA.jl:50
        self.tau_old = self.tau.copy()
        self.C = np.array([self.counts[self.off[d]:self.off[d + 1]].sum() for d in range(self.M)])

    def _sl(self, d):
        return slice(self.off[d], self.off[d + 1])

    @staticmethod
    def _additive_logistic(x):                                             # utils.jl:114-121, dims = the topic axis
        x = np.exp(x - x.max(axis=1, keepdims=True))
        return x / x.sum(axis=1, keepdims=True)

    # fLDA.jl:198-201
    def update_phi(self, d):
        terms, _ = self._doc(d)
        self.phi = self._additive_logistic(self.tau[self._sl(d)][:, None] * np.log(self.beta[terms] + EPSILON) + self.Elogtheta[d][None, :])

    # fLDA.jl:189-194
    def update_tau(self, d):
        terms, _ = self._doc(d)
        s = self._sl(d)
        self.tau_old[s] = self.tau[s]
        with np.errstate(divide="ignore", over="ignore"):
            pr = np.prod(self.beta[terms] ** (-self.phi), axis=1)
            self.tau[s] = self.eta / ((self.eta + (1 - self.eta) * (self.kappa[terms] * pr)) + EPSILON)

    # fLDA.jl:182-185
    def update_gamma(self, d):
        _, counts = self._doc(d)
        self.gamma[d] = EPSILON + (self.alpha + counts @ self.phi)

    # fLDA.jl:168-171, 156-159
    def update_beta_doc(self, d):
        terms, counts = self._doc(d)
        t = self.tau[self._sl(d)]
        self.beta_temp[terms] += self.phi * (t * counts)[:, None]
        self.kappa_temp[terms] += (1 - t) * counts

    # fLDA.jl:149-153
    def update_kappa(self):
        self.kappa_old = self.kappa
        self.kappa = self.kappa_temp / self.kappa_temp.sum()
        self.kappa_temp = np.zeros(self.V)

    # fLDA.jl:119-121
    def update_eta(self):
        self.eta = sum(np.dot(self.tau[self._sl(d)], self.counts[self._sl(d)]) for d in range(self.M)) / self.C.sum()

    # fLDA.jl:62-117
    def update_elbo(self):
        elbo = 0.0
        a = self.alpha
        for d in range(self.M):
            terms, counts = self._doc(d)
            s = self._sl(d)
            phi = self._additive_logistic(self.tau_old[s][:, None] * np.log(self.beta_old[terms] + EPSILON) + self.Elogtheta_old[d][None, :])
            Et, g, tau = self.Elogtheta[d], self.gamma[d], self.tau[s]
            Elogptheta = _finite(gammaln(a.sum())) - _finite(gammaln(a).sum()) + np.dot(a - 1, Et)
            tc = np.dot(tau, counts)
            Elogpc = np.log(self.eta**tc * (1 - self.eta) ** (self.C[d] - tc) + EPSILON)
            Elogpz = np.dot(counts @ phi, Et)
            Elogpw = np.sum((phi * np.log(self.beta[terms] + EPSILON)) * (counts * tau)[:, None]) + np.dot(counts * (1 - tau), np.log(self.kappa[terms] + EPSILON))
            if self.K == 1:
                ent_dir = 0.0
            else:  # utils.jl:163-180
                ent_dir = gammaln(g).sum() - gammaln(g.sum()) + (g.sum() - self.K) * digamma(g.sum()) - np.dot(g - 1, digamma(g))
            p0 = 1 - tau
            with np.errstate(divide="ignore", invalid="ignore"):
                hb = np.where((p0 == 0) | (p0 == 1), 0.0, -(p0 * np.log(p0) + tau * np.log(tau)))
                hz = -np.sum(np.where(phi > 0, phi * np.log(phi), 0.0), axis=1)
            elbo += Elogptheta + Elogpc + Elogpz + Elogpw + ent_dir + np.dot(counts, hb) + np.dot(counts, hz)
        self.elbo = elbo
        return elbo

    def update_elbo_device_form(self):
        """fLDA.jl:62-117 without a logarithm per (token, topic): the decomposition a fused device ELBO can evaluate from what the
        E-step's scatter pass already holds (cf. LDATwin.update_elbo_device_form).  Valid after an E-step + M-step +
        alpha / eta updates.  In base 2 as on the device: L = log2(beta_old + eps), x_ni = tau_old_n L_ni + (Elogtheta_old_i - mx) log2 e,
        p = 2^x, s_n = sum_i p_ni, q_n = sum_i p_ni L_ni, phi = p / s, g_i = sum_n c_n phi_ni:
          sum_n c_n H(phi_n) = sum_n c_n (ln s_n - ln 2 tau_old_n q_n / s_n) - sum_i (Elogtheta_old_i - mx) g_i
          Elogpz             = g . Elogtheta
          H(Dirichlet)       = sum_i lnG(gamma_i) - lnG(sum gamma) - sum_i (gamma_i - 1) Elogtheta_i          (K > 1)
          Elogpw             = sum_iw S_iw ln(beta_iw + eps) + sum_w KS_w ln(kappa_w + eps),   S, KS = the statistics of the E-step
          Elogptheta         = M (lnG(sum alpha) - sum lnG(alpha)) + (alpha - 1) . sum_d Elogtheta_d
        Elogpc (saturating at ln eps for long documents) and the Bernoulli entropy of tau stay per document / per token."""
        a, K = self.alpha, self.K
        S, KS = np.zeros((self.V, K)), np.zeros(self.V)
        docs = 0.0
        ln2 = np.log(2.0)
        for d in range(self.M):
            terms, counts = self._doc(d)
            sl = self._sl(d)
            tau, tauo = self.tau[sl], self.tau_old[sl]
            Eo, Et, gam = self.Elogtheta_old[d], self.Elogtheta[d], self.gamma[d]
            mx = Eo.max()
            L = np.log2(self.beta_old[terms] + EPSILON)
            p = np.exp2(tauo[:, None] * L + ((Eo - mx) / ln2)[None, :])
            sn = p.sum(axis=1)
            qn = (p * L).sum(axis=1)
            phi = p / sn[:, None]
            g = counts @ phi
            S[terms] += phi * (counts * tau)[:, None]
            KS[terms] += (1 - tau) * counts
            docs += np.dot(counts, np.log(sn) - ln2 * tauo * qn / sn) - np.dot(Eo - mx, g)            # sum_n c_n H(phi_n)
            docs += np.dot(g, Et)                                                                     # Elogpz
            if K > 1:
                docs += gammaln(gam).sum() - gammaln(gam.sum()) - np.dot(gam - 1, Et)                  # H(Dirichlet(gamma_d))
            tc = np.dot(tau, counts)
            docs += np.log(self.eta**tc * (1 - self.eta) ** (self.C[d] - tc) + EPSILON)               # Elogpc
            p0 = 1 - tau
            with np.errstate(divide="ignore", invalid="ignore"):
                docs += np.dot(counts, np.where((p0 == 0) | (p0 == 1), 0.0, -(p0 * np.log(p0) + tau * np.log(tau))))
        glob = self.M * (_finite(gammaln(a.sum())) - _finite(gammaln(a).sum())) + np.dot(a - 1, self.Elogtheta.sum(axis=0))
        glob += np.sum(S * np.log(self.beta + EPSILON)) + np.dot(KS, np.log(self.kappa + EPSILON))
        return docs + glob

    # fLDA.jl:214-247
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
                    self.update_tau(d)
                    self.update_gamma(d)
                    self.update_Elogtheta(d)
                    if np.linalg.norm(self.Elogtheta[d] - self.Elogtheta_old[d]) < vtol:
                        break
                self.update_beta_doc(d)
            self.update_beta()
            self.update_kappa()
            self.update_alpha(niter, ntol)
            self.update_eta()
            if k % checkelbo == 0:
                old = self.elbo
                delta = self.update_elbo() - old
                trace[k] = self.elbo
                if delta < tol:
                    break
        return np.array(trace)


class FCTMTwin(CTMTwin):
    """src/fCTM.jl, line for line (filtered CTM: CTM + per-token tau, corpus distribution kappa; eta stays at its initial value,
    update_eta! being commented out of the training loop, fCTM.jl:279)."""

    def __init__(self, N_cumsum, terms, counts, K, V, beta, kappa, eta=0.5):
        super().__init__(N_cumsum, terms, counts, K, V, beta)
        self.eta = float(eta)
        self.kappa = np.array(kappa, dtype=np.float64)
        self.kappa_old = self.kappa.copy()
        self.kappa_temp = np.zeros(V)
        self.tau = np.full(len(self.terms), self.eta)
        self.tau_old = self.tau.copy()

    def _sl(self, d):
        return slice(self.off[d], self.off[d + 1])

    def update_phi(self, d):  # fCTM.jl:239-242
        terms, _ = self._doc(d)
        self.phi = self._softmax_rows(self.tau[self._sl(d)][:, None] * np.log(self.beta[terms] + EPSILON) + self.lam[d][None, :])

    def update_tau(self, d):  # fCTM.jl:230-235
        terms, _ = self._doc(d)
        s = self._sl(d)
        self.tau_old[s] = self.tau[s]
        with np.errstate(divide="ignore", over="ignore"):
            pr = np.prod(self.beta[terms] ** (-self.phi), axis=1)
            self.tau[s] = self.eta / ((self.eta + (1 - self.eta) * (self.kappa[terms] * pr)) + EPSILON)

    def update_elbo(self):  # fCTM.jl:67-130
        elbo = 0.0
        K = self.K
        _, logdet = np.linalg.slogdet(self.invsigma)
        for d in range(self.M):
            terms, counts = self._doc(d)
            s = self._sl(d)
            phi = self._softmax_rows(self.tau_old[s][:, None] * np.log(self.beta_old[terms] + EPSILON) + self.lam_old[d][None, :])
            lam, v, tau = self.lam[d], self.vsq[d], self.tau[s]
            df = lam - self.mu
            Elogpeta = 0.5 * (logdet - K * np.log(2 * np.pi) - np.dot(np.diag(self.invsigma), v) - df @ self.invsigma @ df)
            tc = np.dot(tau, counts)
            Elogpc = np.log(self.eta**tc * (1 - self.eta) ** (self.C[d] - tc) + EPSILON)
            Elogpz = np.dot(phi @ lam, counts) - self.C[d] * (np.exp(lam + 0.5 * v - self.logzeta[d]).sum() + self.logzeta[d] - 1)
            Elogpw = np.sum((phi * np.log(self.beta[terms] + EPSILON)) * (counts * tau)[:, None]) + np.dot(counts * (1 - tau), np.log(self.kappa[terms] + EPSILON))
            ent_eta = 0.5 * (K * (np.log(2 * np.pi) + 1) + np.log(v).sum())
            p0 = 1 - tau
            with np.errstate(divide="ignore", invalid="ignore"):
                hb = np.where((p0 == 0) | (p0 == 1), 0.0, -(p0 * np.log(p0) + tau * np.log(tau)))
                plogp = np.where(phi > 0, phi * np.log(phi), 0.0)
            elbo += Elogpeta + Elogpc + Elogpz + Elogpw + ent_eta + np.dot(counts, hb) - (plogp.sum(axis=1) * counts).sum()
        self.elbo = elbo
        return elbo

    def update_elbo_device_form(self):
        """fCTM.jl:67-130 without a logarithm per (token, topic) (cf. FLDATwin / CTMTwin.update_elbo_device_form).  In base 2 over
        L = log2(beta_old + eps): x_ni = tau_old_n L_ni + (lambda_old_i - mx) log2 e, p = 2^x, s = sum_i p, q = sum_i p L, phi = p / s,
        g_i = sum_n c_n phi_ni:
          sum_n c_n H(phi_n) = sum_n c_n (ln s_n - ln 2 tau_old_n q_n / s_n) - sum_i (lambda_old_i - mx) g_i
          Elogpz             = g . lambda - C_d (sum_i exp(lambda_i + v_i / 2 - ln zeta) + ln zeta - 1)
          Elogpw             = sum_iw S_iw ln(beta_iw + eps) + sum_w KS_w ln(kappa_w + eps)        (the statistics of the E-step)
        Elogpeta, the Gaussian entropy, Elogpc and the Bernoulli entropy of tau are K-vector / per-token algebra."""
        K = self.K
        _, logdet = np.linalg.slogdet(self.invsigma)
        S, KS = np.zeros((self.V, K)), np.zeros(self.V)
        docs = 0.0
        ln2 = np.log(2.0)
        for d in range(self.M):
            terms, counts = self._doc(d)
            sl = self._sl(d)
            tau, tauo = self.tau[sl], self.tau_old[sl]
            lo, lam, v = self.lam_old[d], self.lam[d], self.vsq[d]
            mx = lo.max()
            L = np.log2(self.beta_old[terms] + EPSILON)
            p = np.exp2(tauo[:, None] * L + ((lo - mx) / ln2)[None, :])
            sn = p.sum(axis=1)
            qn = (p * L).sum(axis=1)
            phi = p / sn[:, None]
            g = counts @ phi
            S[terms] += phi * (counts * tau)[:, None]
            KS[terms] += (1 - tau) * counts
            df = lam - self.mu
            docs += 0.5 * (logdet - K * np.log(2 * np.pi) - np.dot(np.diag(self.invsigma), v) - df @ self.invsigma @ df)
            docs += np.dot(g, lam) - self.C[d] * (np.exp(lam + 0.5 * v - self.logzeta[d]).sum() + self.logzeta[d] - 1)
            docs += 0.5 * (K * (np.log(2 * np.pi) + 1) + np.log(v).sum())
            docs += np.dot(counts, np.log(sn) - ln2 * tauo * qn / sn) - np.dot(lo - mx, g)
            tc = np.dot(tau, counts)
            docs += np.log(self.eta**tc * (1 - self.eta) ** (self.C[d] - tc) + EPSILON)
            p0 = 1 - tau
            with np.errstate(divide="ignore", invalid="ignore"):
                docs += np.dot(counts, np.where((p0 == 0) | (p0 == 1), 0.0, -(p0 * np.log(p0) + tau * np.log(tau))))
        return docs + np.sum(S * np.log(self.beta + EPSILON)) + np.dot(KS, np.log(self.kappa + EPSILON))

    def train(self, iter=150, tol=1.0, niter=1000, ntol=None, viter=10, vtol=None, checkelbo=1):  # fCTM.jl:249-290
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
                    self.update_tau(d)
                    self.update_logzeta(d)
                    self.update_lambda(d, niter, ntol)
                    self.update_vsq(d, niter, ntol)
                    if np.linalg.norm(self.lam[d] - self.lam_old[d]) < vtol:
                        break
                t = self.tau[self._sl(d)]
                self.beta_temp[terms] += self.phi * (t * counts)[:, None]      # fCTM.jl:175-178
                self.kappa_temp[terms] += (1 - t) * counts                     # fCTM.jl:162-165
            self.beta_old = self.beta
            self.beta = self.beta_temp / self.beta_temp.sum(axis=0, keepdims=True)
            self.beta_temp = np.zeros((self.V, K))
            self.kappa_old = self.kappa
            self.kappa = self.kappa_temp / self.kappa_temp.sum()
            self.kappa_temp = np.zeros(self.V)
            dl = self.lam - self.mu[None, :]                                   # update_sigma! with the old mu
            self.sigma = (np.diag(self.vsq.sum(axis=0)) + dl.T @ dl) / self.M
            self.invsigma = np.linalg.inv(self.sigma)
            self.mu = self.lam.sum(axis=0) / self.M
            if k % checkelbo == 0:
                old = self.elbo
                delta = self.update_elbo() - old
                trace[k] = self.elbo
                if delta < tol:
                    break
        return np.array(trace)
