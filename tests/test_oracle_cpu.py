"""CPU tests of the oracle (test infrastructure): special functions vs SciPy, the C restatement vs the
independent NumPy twin, the committed golden vectors, invariants of check_model (modelutils.jl:39-67)."""
import os

import numpy as np
import pytest
from scipy.special import digamma, gammaln, polygamma

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_special_functions_match_scipy(orc):
    lib = orc.load()
    xs = np.concatenate([np.geomspace(1e-3, 1e4, 400), np.linspace(0.5, 12, 47), [1.0, 2.0, 6.0, 10.0]])
    d = np.array([lib.orc_digamma_export(float(x)) for x in xs])
    t = np.array([lib.orc_trigamma_export(float(x)) for x in xs])
    g = np.array([lib.orc_lgamma_export(float(x)) for x in xs])
    np.testing.assert_allclose(d, digamma(xs), rtol=4e-15, atol=4e-15)
    np.testing.assert_allclose(t, polygamma(1, xs), rtol=4e-15)
    np.testing.assert_allclose(g, gammaln(xs), rtol=1e-14, atol=1e-14)
    assert abs(lib.orc_digamma_export(1.0) + np.euler_gamma) < 1e-15          # psi(1) = -gamma (LDA.jl:38)


def test_epsilon_is_eps_of_1e_minus_14():
    from oracle.numpy_twin import EPSILON
    assert EPSILON == np.spacing(1e-14) == 2.0 ** -99                           # utils.jl:3


@pytest.mark.parametrize("K,M,V,seed", [(5, 100, 500, 0), (3, 40, 120, 4), (1, 20, 60, 5), (12, 30, 200, 6)])
def test_c_oracle_matches_numpy_twin(orc, K, M, V, seed):
    import topicmodelsvb_b200.synth as synth
    from oracle.numpy_twin import LDATwin

    c = synth.gencorp_lda(M=M, V=V, K=max(K, 2), seed=seed)
    beta0 = synth.init_beta(K, c.V, seed=7)
    st = orc.LDAState(K, c.M, c.V, beta=beta0)
    tr, sw, done = orc.lda_train(st, c.N_cumsum, c.terms, c.counts, iter=6, tol=-np.inf)
    tw = LDATwin(c.N_cumsum, c.terms, c.counts, K, c.V, beta0)
    t2 = tw.train(iter=6, tol=-np.inf)
    np.testing.assert_allclose(tr, t2, rtol=1e-12)
    np.testing.assert_allclose(st.alpha, tw.alpha, rtol=1e-10)
    np.testing.assert_allclose(st.beta, tw.beta, rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(st.gamma, tw.gamma, rtol=1e-10)
    np.testing.assert_allclose(st.Elogtheta, tw.Elogtheta, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("K,M,V,seed", [(5, 100, 500, 0), (1, 20, 60, 5), (12, 30, 200, 6)])
def test_device_elbo_decomposition_equals_literal_elbo(K, M, V, seed):
    """The decomposition the CUDA path evaluates (no logarithm per token and topic) equals
    update_elbo! (LDA.jl:50-93) after every outer iteration, in fp64 on the CPU."""
    import topicmodelsvb_b200.synth as synth
    from oracle.numpy_twin import LDATwin

    c = synth.gencorp_lda(M=M, V=V, K=max(K, 2), seed=seed)
    tw = LDATwin(c.N_cumsum, c.terms, c.counts, K, c.V, synth.init_beta(K, c.V, seed=7))
    for it in range(4):
        tr = tw.train(iter=1, tol=-np.inf)
        lit, dev = tr[1], tw.update_elbo_device_form()
        assert abs(dev - lit) <= 1e-11 * abs(lit), (it, lit, dev)


@pytest.mark.parametrize("K,M,V,seed", [(4, 40, 150, 2), (1, 15, 50, 3), (7, 25, 120, 8)])
def test_ctm_elbo_decomposition_equals_literal_elbo(K, M, V, seed):
    """The logarithm-free form of the CTM ELBO (CTMTwin.update_elbo_device_form) equals update_elbo! (CTM.jl:56-98)."""
    import topicmodelsvb_b200.synth as synth
    from oracle.numpy_twin import CTMTwin

    c = synth.gencorp_lda(M=M, V=V, K=max(K, 2), seed=seed)
    tw = CTMTwin(c.N_cumsum, c.terms, c.counts, K, c.V, synth.init_beta(K, c.V, seed=7))
    for it in range(3):
        tr = tw.train(iter=1, tol=-np.inf)
        lit, dev = tr[1], tw.update_elbo_device_form()
        assert abs(dev - lit) <= 1e-11 * abs(lit), (it, lit, dev)


@pytest.mark.parametrize("K,seed", [(4, 0), (1, 1), (6, 2)])
def test_ctpf_entropy_decomposition_equals_literal_elbo(K, seed):
    """The logarithm-free form of the two CTPF entropy sums (CTPFTwin.update_elbo_device_form) equals update_elbo!."""
    import topicmodelsvb_b200.synth as synth
    from oracle.numpy_twin import CTPFTwin

    c = synth.gencorp_ctpf(M=40, V=150, U=25, K=3, seed=seed)
    tw = CTPFTwin(c.N_cumsum, c.terms, c.counts, c.R_cumsum, c.readers, c.ratings, K, c.V, c.U, synth.init_alef(K, c.V, seed=7))
    for it in range(3):
        tr = tw.train(iter=1, tol=-np.inf)
        lit, dev = tr[1], tw.update_elbo_device_form()
        assert abs(dev - lit) <= 1e-11 * abs(lit), (it, lit, dev)


def test_golden_lda_cfg0(orc):
    g = np.load(os.path.join(GOLD, "lda_cfg0.npz"))
    K, V = int(g["K"]), int(g["V"])
    M = len(g["N_cumsum"]) - 1
    st = orc.LDAState(K, M, V, beta=g["beta0"])
    tr, sw, done = orc.lda_train(st, g["N_cumsum"], g["terms"].astype(np.int64), g["counts"].astype(np.int64), iter=20, tol=0.0)
    np.testing.assert_allclose(tr, g["elbo"], rtol=1e-12)
    np.testing.assert_array_equal(sw, g["sweeps"])
    np.testing.assert_allclose(st.alpha, g["alpha"], rtol=1e-10)
    np.testing.assert_allclose(st.beta, g["beta"], rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(st.gamma, g["gamma"], rtol=1e-10)
    # ELBO never decreases on this configuration (CAVI ascent; the alpha barrier step is benign here)
    assert np.all(np.diff(tr) > 0)
    # invariants of check_model(::LDA) (modelutils.jl:39-67)
    np.testing.assert_allclose(st.beta.sum(axis=0), 1.0, rtol=1e-12)
    assert np.all(st.alpha > 0) and np.all(st.gamma > 0) and np.all(st.Elogtheta <= 0)


def test_threaded_oracle_only_reassociates(orc):
    import topicmodelsvb_b200.synth as synth

    c = synth.gencorp_lda(M=150, V=300, K=4, seed=8)
    beta0 = synth.init_beta(6, c.V, seed=7)
    a = orc.LDAState(6, c.M, c.V, beta=beta0)
    b = orc.LDAState(6, c.M, c.V, beta=beta0)
    ta, _, _ = orc.lda_train(a, c.N_cumsum, c.terms, c.counts, iter=5, tol=0.0, nthreads=1)
    tb, _, _ = orc.lda_train(b, c.N_cumsum, c.terms, c.counts, iter=5, tol=0.0, nthreads=4)
    np.testing.assert_allclose(ta, tb, rtol=1e-12)
    np.testing.assert_allclose(a.beta, b.beta, rtol=1e-9, atol=1e-300)


def test_checkelbo_and_tol_semantics(orc):
    """check_elbo! (modelutils.jl:574-585): evaluated when k % checkelbo == 0; delta < tol stops."""
    import topicmodelsvb_b200.synth as synth

    c = synth.gencorp_lda(M=60, V=200, K=3, seed=1)
    beta0 = synth.init_beta(4, c.V, seed=7)
    st = orc.LDAState(4, c.M, c.V, beta=beta0)
    tr, _, done = orc.lda_train(st, c.N_cumsum, c.terms, c.counts, iter=6, tol=0.0, checkelbo=2)
    assert done == 6 and np.isfinite(tr[[0, 2, 4, 6]]).all() and np.isnan(tr[[1, 3, 5]]).all()
    st = orc.LDAState(4, c.M, c.V, beta=beta0)
    tr, _, done = orc.lda_train(st, c.N_cumsum, c.terms, c.counts, iter=50, tol=1e12)
    assert done == 1
    st = orc.LDAState(4, c.M, c.V, beta=beta0)
    tr, _, done = orc.lda_train(st, c.N_cumsum, c.terms, c.counts, iter=3, tol=0.0, checkelbo=float("inf"))
    assert done == 3 and np.isnan(tr).all()


def test_estep_scatter_equals_train_statistics(orc):
    """orc.lda_estep (what bench.py's cpu_baseline times) == the E-step inside train!."""
    import topicmodelsvb_b200.synth as synth

    c = synth.gencorp_lda(M=50, V=150, K=3, seed=2)
    K = 5
    beta0 = synth.init_beta(K, c.V, seed=7)
    a = orc.LDAState(K, c.M, c.V, beta=beta0)
    stats, sweeps = orc.lda_estep(a, c.N_cumsum, c.terms, c.counts)
    b = orc.LDAState(K, c.M, c.V, beta=beta0)
    _, sw, _ = orc.lda_train(b, c.N_cumsum, c.terms, c.counts, iter=1, tol=0.0, checkelbo=float("inf"))
    assert sweeps == int(sw[0])
    np.testing.assert_allclose(stats / stats.sum(axis=0, keepdims=True), b.beta, rtol=1e-12)
    np.testing.assert_allclose(stats.sum(), c.counts.sum(), rtol=1e-12)       # every token's phi sums to one
    phi = orc.lda_phi(K, c.M, c.N_cumsum, c.terms, b.beta_old, b.Elogtheta_old)
    np.testing.assert_allclose(phi.sum(axis=1), 1.0, rtol=1e-12)


@pytest.mark.parametrize("K,seed", [(6, 1), (1, 2), (11, 3)])
def test_ctm_c_oracle_matches_numpy_twin(orc, K, seed):
    import topicmodelsvb_b200.synth as synth
    from oracle.numpy_twin import CTMTwin

    c = synth.gencorp_lda(M=40, V=150, K=4, seed=seed)
    beta0 = synth.init_beta(K, c.V, seed=7)
    st = orc.CTMState(K, c.M, c.V, beta0)
    tr, sw, done = orc.ctm_train(st, c.N_cumsum, c.terms, c.counts, iter=4, tol=0.0)
    tw = CTMTwin(c.N_cumsum, c.terms, c.counts, K, c.V, beta0)
    t2 = tw.train(iter=4, tol=0.0)
    np.testing.assert_allclose(tr, t2, rtol=1e-10)
    np.testing.assert_allclose(st.lam, tw.lam, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(st.vsq, tw.vsq, rtol=1e-8)
    np.testing.assert_allclose(st.sigma, tw.sigma, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(st.mu, tw.mu, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(st.invsigma @ st.sigma, np.eye(K), atol=1e-9)          # CTM.jl:110
    assert np.all(st.vsq > 0) and np.all(np.linalg.eigvalsh(st.sigma) > 0)             # check_model, modelutils.jl:113-119


@pytest.mark.parametrize("K,ratings_max", [(5, 1), (3, 4), (1, 2)])
def test_ctpf_c_oracle_long_form_matches_closed_form_twin(orc, K, ratings_max):
    """The C oracle evaluates the reference's Binomial x lnGamma sums and Distributions' entropy(Multinomial) literally
    (CTPF.jl:111-141,179-194); the NumPy twin uses the closed form they collapse to.  Agreement pins that identity."""
    import topicmodelsvb_b200.synth as synth
    from oracle.numpy_twin import CTPFTwin

    c = synth.gencorp_ctpf(M=30, V=120, U=25, K=3, seed=4)
    if ratings_max > 1:
        c = c._replace(ratings=np.random.default_rng(5).integers(1, ratings_max + 1, size=len(c.readers)))
    alef0 = synth.init_alef(K, c.V, seed=7)
    st = orc.CTPFState(K, c.M, c.V, c.U, alef0)
    tr, sw, done = orc.ctpf_train(st, c, iter=4, tol=0.0)
    tw = CTPFTwin(c.N_cumsum, c.terms, c.counts, c.R_cumsum, c.readers, c.ratings, K, c.V, c.U, alef0)
    t2 = tw.train(iter=4, tol=0.0)
    n = min(int(np.isfinite(tr).sum()), int(np.isfinite(t2).sum()))
    np.testing.assert_allclose(tr[:n], t2[:n], rtol=1e-11)
    np.testing.assert_allclose(st.gimel, tw.gimel, rtol=1e-9)
    np.testing.assert_allclose(st.zayin, tw.zayin, rtol=1e-9)
    np.testing.assert_allclose(st.alef, tw.alef, rtol=1e-9)
    np.testing.assert_allclose(st.he, tw.he, rtol=1e-9)
    for nme in ("bet", "vav", "dalet", "het"):
        np.testing.assert_allclose(getattr(st, nme), getattr(tw, nme), rtol=1e-10)


def test_golden_ctm_and_ctpf(orc):
    g = np.load(os.path.join(GOLD, "ctm_cfg.npz"))
    K, V = int(g["K"]), int(g["V"])
    M = len(g["N_cumsum"]) - 1
    st = orc.CTMState(K, M, V, g["beta0"])
    tr, _, _ = orc.ctm_train(st, g["N_cumsum"], g["terms"].astype(np.int64), g["counts"].astype(np.int64), iter=len(g["elbo"]) - 1, tol=0.0)
    np.testing.assert_allclose(tr, g["elbo"], rtol=1e-10)
    np.testing.assert_allclose(st.sigma, g["sigma"], rtol=1e-8, atol=1e-12)
    import topicmodelsvb_b200.synth as synth
    g = np.load(os.path.join(GOLD, "ctpf_cfg.npz"))
    K, V, U = int(g["K"]), int(g["V"]), int(g["U"])
    c = synth.CSR(len(g["N_cumsum"]) - 1, V, g["N_cumsum"], g["terms"].astype(np.int64), g["counts"].astype(np.int64), U,
                  g["R_cumsum"], g["readers"].astype(np.int64), g["ratings"].astype(np.int64))
    st = orc.CTPFState(K, c.M, V, U, g["alef0"])
    tr, _, _ = orc.ctpf_train(st, c, iter=len(g["elbo"]) - 1, tol=0.0)
    np.testing.assert_allclose(tr, g["elbo"], rtol=1e-10)
    np.testing.assert_allclose(st.vav, g["vav"], rtol=1e-10)


def test_flda_c_oracle_equals_numpy_twin(orc):
    """oracle/flda_oracle.c == the literal NumPy transcription of src/fLDA.jl (FLDATwin) to <= 1e-12 relative: ELBO trajectory,
    eta, tau, beta, kappa, alpha; the OpenMP document split only re-associates fp64 sums."""
    import topicmodelsvb_b200 as tm
    from oracle.numpy_twin import FLDATwin

    c = tm.synth.gencorp_lda(M=40, V=120, K=4, seed=3)
    for K in (1, 5):
        beta0 = tm.synth.init_beta(K, c.V, seed=7)
        kappa = np.random.default_rng(1).dirichlet(np.ones(c.V))
        tw = FLDATwin(c.N_cumsum, c.terms, c.counts, K, c.V, beta0, kappa)
        t1 = tw.train(iter=4, tol=0.0)
        st = orc.FLDAState(K, c.M, c.V, len(c.terms), beta0, kappa)
        t2, sweeps, done = orc.flda_train(st, c.N_cumsum, c.terms, c.counts, iter=4, tol=0.0)
        assert done == 4 and np.all(sweeps > 0)
        np.testing.assert_allclose(t2, t1, rtol=1e-12)
        assert abs(st.eta[0] - tw.eta) < 1e-13
        np.testing.assert_allclose(st.tau, tw.tau, rtol=1e-11, atol=1e-15)
        np.testing.assert_allclose(st.tau_old, tw.tau_old, rtol=1e-11, atol=1e-15)
        np.testing.assert_allclose(st.beta, tw.beta, rtol=1e-11, atol=1e-300)
        np.testing.assert_allclose(st.kappa, tw.kappa, rtol=1e-11)
        np.testing.assert_allclose(st.alpha, tw.alpha, rtol=1e-11)
        np.testing.assert_allclose(st.gamma, tw.gamma, rtol=1e-11)
        st4 = orc.FLDAState(K, c.M, c.V, len(c.terms), beta0, kappa)
        t4, _, _ = orc.flda_train(st4, c.N_cumsum, c.terms, c.counts, iter=4, tol=0.0, nthreads=4)
        np.testing.assert_allclose(t4, t2, rtol=1e-12)
    # checkelbo = 2 evaluates every other iteration; a large tol stops at the first check
    st = orc.FLDAState(5, c.M, c.V, len(c.terms), tm.synth.init_beta(5, c.V, seed=7), kappa)
    t, _, done = orc.flda_train(st, c.N_cumsum, c.terms, c.counts, iter=6, tol=1e12, checkelbo=2)
    assert done == 2 and np.isfinite(t[[0, 2]]).all() and np.isnan(t[1])


def test_fctm_c_oracle_equals_numpy_twin(orc):
    """oracle/fctm_oracle.c == the literal NumPy transcription of src/fCTM.jl (FCTMTwin)."""
    import topicmodelsvb_b200 as tm
    from oracle.numpy_twin import FCTMTwin

    c = tm.synth.gencorp_lda(M=30, V=100, K=4, seed=3)
    kappa = np.random.default_rng(1).dirichlet(np.ones(c.V))
    for K in (1, 4):
        beta0 = tm.synth.init_beta(K, c.V, seed=7)
        tw = FCTMTwin(c.N_cumsum, c.terms, c.counts, K, c.V, beta0, kappa)
        t1 = tw.train(iter=3, tol=0.0)
        st = orc.FCTMState(K, c.M, c.V, len(c.terms), beta0, kappa)
        t2, sweeps, done = orc.fctm_train(st, c.N_cumsum, c.terms, c.counts, iter=3, tol=0.0)
        assert done == 3
        np.testing.assert_allclose(t2, t1, rtol=1e-11)
        np.testing.assert_allclose(st.tau, tw.tau, rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(st.lam, tw.lam, rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(st.vsq, tw.vsq, rtol=1e-8)
        np.testing.assert_allclose(st.kappa, tw.kappa, rtol=1e-9)
        np.testing.assert_allclose(st.sigma, tw.sigma, rtol=1e-8, atol=1e-12)
        st4 = orc.FCTMState(K, c.M, c.V, len(c.terms), beta0, kappa)
        t4, _, _ = orc.fctm_train(st4, c.N_cumsum, c.terms, c.counts, iter=3, tol=0.0, nthreads=4)
        np.testing.assert_allclose(t4, t2, rtol=1e-11)


@pytest.mark.parametrize("K,M,V,seed", [(5, 40, 150, 2), (1, 15, 60, 3), (8, 30, 120, 5)])
def test_flda_elbo_device_form_equals_literal_form(K, M, V, seed):
    """The logarithm-free form of the fLDA ELBO (FLDATwin.update_elbo_device_form: what a fused device ELBO would evaluate)
    equals update_elbo! (fLDA.jl:62-117) after every outer iteration, in fp64 on the CPU."""
    import topicmodelsvb_b200.synth as synth
    from oracle.numpy_twin import FLDATwin

    c = synth.gencorp_lda(M=M, V=V, K=max(K, 2), seed=seed)
    kappa0 = np.random.default_rng(seed).dirichlet(np.ones(c.V))
    tw = FLDATwin(c.N_cumsum, c.terms, c.counts, K, c.V, synth.init_beta(K, c.V, seed=7), kappa0)
    for it in range(4):
        tr = tw.train(iter=1, tol=-np.inf)
        lit, dev = tr[1], tw.update_elbo_device_form()
        assert abs(dev - lit) <= 1e-11 * abs(lit), (it, lit, dev)


@pytest.mark.parametrize("K,M,V,seed", [(4, 30, 120, 2), (1, 12, 50, 3), (6, 25, 100, 7)])
def test_fctm_elbo_device_form_equals_literal_form(K, M, V, seed):
    """The same for fCTM (FCTMTwin.update_elbo_device_form vs update_elbo!, fCTM.jl:67-130)."""
    import topicmodelsvb_b200.synth as synth
    from oracle.numpy_twin import FCTMTwin

    c = synth.gencorp_lda(M=M, V=V, K=max(K, 2), seed=seed)
    kappa0 = np.random.default_rng(seed).dirichlet(np.ones(c.V))
    tw = FCTMTwin(c.N_cumsum, c.terms, c.counts, K, c.V, synth.init_beta(K, c.V, seed=7), kappa0)
    for it in range(3):
        tr = tw.train(iter=1, tol=-np.inf)
        lit, dev = tr[1], tw.update_elbo_device_form()
        assert abs(dev - lit) <= 1e-11 * abs(lit), (it, lit, dev)
