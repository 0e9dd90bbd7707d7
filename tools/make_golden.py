"""Generates tests/golden/*.npz from the CPU oracle (both restatements must agree first).

The reference ships no golden vectors and cannot run here (no Julia), so these fixtures pin OUR
restatement against regressions and give the GPU tests committed reference trajectories; they do not
pin parity with the reference itself ("parity unpinned").

usage: python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle.numpy_twin import CTMTwin, CTPFTwin, FCTMTwin, FLDATwin, LDATwin  # noqa: E402
import topicmodelsvb_b200.synth as synth  # noqa: E402


def lda_cfg0():
    """M=100, V=500, K=5, iter=20, tol=0, viter=10, vtol=ntol=1/K^2."""
    K = 5
    c = synth.gencorp_lda(M=100, V=500, K=K, seed=0)
    beta0 = synth.init_beta(K, c.V, seed=7)
    st = oracle.LDAState(K, c.M, c.V, beta=beta0)
    trace, sweeps, done = oracle.lda_train(st, c.N_cumsum, c.terms, c.counts, iter=20, tol=0.0)
    tw = LDATwin(c.N_cumsum, c.terms, c.counts, K, c.V, beta0)
    t2 = tw.train(iter=20, tol=0.0)
    assert np.max(np.abs(trace - t2) / np.abs(t2)) < 1e-12
    assert np.allclose(st.beta, tw.beta, rtol=1e-10, atol=1e-300)
    return dict(K=K, V=c.V, N_cumsum=c.N_cumsum, terms=c.terms.astype(np.int32), counts=c.counts.astype(np.int32),
                beta0=beta0, elbo=trace, sweeps=sweeps, alpha=st.alpha, beta=st.beta, gamma=st.gamma,
                Elogtheta=st.Elogtheta, beta_old=st.beta_old, Elogtheta_old=st.Elogtheta_old)


def ctm_cfg():
    K = 6
    c = synth.gencorp_lda(M=60, V=300, K=4, seed=1)
    beta0 = synth.init_beta(K, c.V, seed=7)
    st = oracle.CTMState(K, c.M, c.V, beta0)
    trace, sweeps, _ = oracle.ctm_train(st, c.N_cumsum, c.terms, c.counts, iter=8, tol=0.0)
    tw = CTMTwin(c.N_cumsum, c.terms, c.counts, K, c.V, beta0)
    assert np.max(np.abs(trace - tw.train(iter=8, tol=0.0)) / np.abs(trace)) < 1e-10
    return dict(K=K, V=c.V, N_cumsum=c.N_cumsum, terms=c.terms.astype(np.int32), counts=c.counts.astype(np.int32), beta0=beta0,
                elbo=trace, sweeps=sweeps, mu=st.mu, sigma=st.sigma, beta=st.beta, lam=st.lam, vsq=st.vsq, logzeta=st.logzeta)


def ctpf_cfg():
    K = 5
    c = synth.gencorp_ctpf(M=60, V=300, U=40, K=4, seed=0)
    alef0 = synth.init_alef(K, c.V, seed=7)
    st = oracle.CTPFState(K, c.M, c.V, c.U, alef0)
    trace, sweeps, _ = oracle.ctpf_train(st, c, iter=8, tol=0.0)
    tw = CTPFTwin(c.N_cumsum, c.terms, c.counts, c.R_cumsum, c.readers, c.ratings, K, c.V, c.U, alef0)
    assert np.nanmax(np.abs(trace - tw.train(iter=8, tol=0.0)) / np.abs(trace)) < 1e-10
    return dict(K=K, V=c.V, U=c.U, N_cumsum=c.N_cumsum, terms=c.terms.astype(np.int32), counts=c.counts.astype(np.int32),
                R_cumsum=c.R_cumsum, readers=c.readers.astype(np.int32), ratings=c.ratings.astype(np.int32), alef0=alef0,
                elbo=trace, sweeps=sweeps, alef=st.alef, he=st.he, bet=st.bet, vav=st.vav, dalet=st.dalet, het=st.het,
                gimel=st.gimel, zayin=st.zayin)


def flda_cfg():
    """Filtered LDA (fLDA.jl): M=80, V=300, K=5, iter=8; kappa0 injected like beta0."""
    K = 5
    c = synth.gencorp_lda(M=80, V=300, K=4, seed=2)
    beta0 = synth.init_beta(K, c.V, seed=7)
    kappa0 = np.random.default_rng(8).dirichlet(np.ones(c.V))
    st = oracle.FLDAState(K, c.M, c.V, len(c.terms), beta0, kappa0)
    trace, sweeps, _ = oracle.flda_train(st, c.N_cumsum, c.terms, c.counts, iter=8, tol=0.0)
    tw = FLDATwin(c.N_cumsum, c.terms, c.counts, K, c.V, beta0, kappa0)
    assert np.max(np.abs(trace - tw.train(iter=8, tol=0.0)) / np.abs(trace)) < 1e-12
    return dict(K=K, V=c.V, N_cumsum=c.N_cumsum, terms=c.terms.astype(np.int32), counts=c.counts.astype(np.int32), beta0=beta0, kappa0=kappa0,
                elbo=trace, sweeps=sweeps, eta=st.eta, alpha=st.alpha, kappa=st.kappa, beta=st.beta, gamma=st.gamma, tau=st.tau)


def fctm_cfg():
    """Filtered CTM (fCTM.jl): M=60, V=300, K=6, iter=6."""
    K = 6
    c = synth.gencorp_lda(M=60, V=300, K=4, seed=3)
    beta0 = synth.init_beta(K, c.V, seed=7)
    kappa0 = np.random.default_rng(8).dirichlet(np.ones(c.V))
    st = oracle.FCTMState(K, c.M, c.V, len(c.terms), beta0, kappa0)
    trace, sweeps, _ = oracle.fctm_train(st, c.N_cumsum, c.terms, c.counts, iter=6, tol=0.0)
    tw = FCTMTwin(c.N_cumsum, c.terms, c.counts, K, c.V, beta0, kappa0)
    assert np.max(np.abs(trace - tw.train(iter=6, tol=0.0)) / np.abs(trace)) < 1e-10
    return dict(K=K, V=c.V, N_cumsum=c.N_cumsum, terms=c.terms.astype(np.int32), counts=c.counts.astype(np.int32), beta0=beta0, kappa0=kappa0,
                elbo=trace, sweeps=sweeps, mu=st.mu, sigma=st.sigma, kappa=st.kappa, beta=st.beta, lam=st.lam, vsq=st.vsq, tau=st.tau)


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    only = set(sys.argv[1:])     # e.g. `python tools/make_golden.py flda_cfg fctm_cfg` leaves the other fixtures untouched
    if only:
        for name in only:
            np.savez_compressed(os.path.join(out, name + ".npz"), **globals()[name]())
        sys.exit(0)
    np.savez_compressed(os.path.join(out, "lda_cfg0.npz"), **lda_cfg0())
    np.savez_compressed(os.path.join(out, "ctm_cfg.npz"), **ctm_cfg())
    np.savez_compressed(os.path.join(out, "ctpf_cfg.npz"), **ctpf_cfg())
    np.savez_compressed(os.path.join(out, "flda_cfg.npz"), **flda_cfg())
    np.savez_compressed(os.path.join(out, "fctm_cfg.npz"), **fctm_cfg())
    for f in sorted(os.listdir(out)):
        print(f, os.path.getsize(os.path.join(out, f)), "bytes")
