"""GPU parity tests for the filtered LDA path (gpufLDA: CUDA through the C ABI vs the fp64 CPU oracle of src/fLDA.jl on the same
seeded inputs).  ELBO tolerance: the north star's 1e-4 relative (asserted 2e-5)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ELBO_RTOL = 2e-5


def _init(tm, c, K, seed=7):
    beta0 = tm.synth.init_beta(K, c.V, seed=seed).astype(np.float32)                 # (V, K)
    kappa0 = np.random.default_rng(seed + 1).dirichlet(np.ones(c.V)).astype(np.float32)
    return beta0, kappa0


def _run_pair(tm, orc, c, K, iters, nthreads=1, **kw):
    beta0, kappa0 = _init(tm, c, K)
    model = tm.gpufLDA(tm.Corpus.from_csr(c), K)
    model.beta = np.array(beta0.T, order="F", copy=True)
    model.kappa = kappa0.copy()
    trace = []
    tm.train(model, iter=iters, tol=0.0, checkelbo=1, printelbo=False, trace=trace, **kw)
    st = orc.FLDAState(K, c.M, c.V, len(c.terms), beta0, kappa0)
    ref, sweeps, done = orc.flda_train(st, c.N_cumsum, c.terms, c.counts, iter=iters, tol=0.0, nthreads=nthreads, **kw)
    return model, np.array(trace), st, ref[np.isfinite(ref)], sweeps


@pytest.mark.parametrize("K", [5, 1, 3, 8, 20, 50, 100])
def test_flda_elbo_trajectory_small(tm, orc, K):
    c = tm.synth.gencorp_lda(M=80, V=300, K=4, seed=K)
    model, trace, st, ref, sweeps = _run_pair(tm, orc, c, K, iters=5)
    if K == 1:
        # one topic: the model converges in the first iteration and the ELBO differences are fp32 noise around zero, so check_elbo!
        # (tol = 0) may stop the device run early; compare the iterations both made
        assert len(trace) >= 3
        ref = ref[: len(trace)]
    assert len(trace) == len(ref)
    np.testing.assert_allclose(trace, ref, rtol=ELBO_RTOL)
    if len(trace) < 6:
        return
    assert abs(model.eta - st.eta[0]) < 1e-4
    np.testing.assert_allclose(model.alpha, st.alpha, rtol=2e-3)
    np.testing.assert_allclose(model.kappa, st.kappa, rtol=5e-3, atol=1e-7)
    np.testing.assert_allclose(model.kappa_old, st.kappa_old, rtol=5e-3, atol=1e-7)
    np.testing.assert_allclose(model.beta.T, st.beta, rtol=1e-2, atol=1e-6)
    np.testing.assert_allclose(model.beta_old.T, st.beta_old, rtol=1e-2, atol=1e-6)
    np.testing.assert_allclose(model.gamma.T, st.gamma, rtol=5e-3, atol=1e-4)
    np.testing.assert_allclose(model.Elogtheta.T, st.Elogtheta, rtol=5e-3, atol=5e-3)
    np.testing.assert_allclose(model.Elogtheta_old.T, st.Elogtheta_old, rtol=5e-3, atol=5e-3)
    np.testing.assert_allclose(model.tau, st.tau, rtol=5e-3, atol=1e-5)
    np.testing.assert_allclose(model.tau_old, st.tau_old, rtol=5e-3, atol=1e-5)
    tm.check_model(model)
    assert abs(model.stats().sweeps - int(sweeps[-1])) <= max(3, 0.02 * sweeps[-1])
    assert np.all((model.tau >= 0) & (model.tau <= 1)) and np.all(model.gamma > 0)
    assert [len(t) for t in model.topics] == [c.V] * K


def test_flda_ragged_and_long_documents(tm, orc):
    """empty documents, one-token documents and documents longer than any shared-memory tile (rows and tau read from L2)"""
    rng = np.random.default_rng(2)
    V, K = 900, 6
    lens = [0, 1, 40, 0, 700, 3, 17, 64, 65, 850]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    terms = np.concatenate([rng.choice(V, n, replace=False) for n in lens]).astype(np.int64)
    counts = rng.integers(1, 5, size=off[-1]).astype(np.int64)
    c = tm.synth.CSR(len(lens), V, off, terms, counts)
    import os
    os.environ["TMVB_TILE_CAP_MAX"] = "128"          # force the overflow path for the two long documents
    try:
        model, trace, st, ref, _ = _run_pair(tm, orc, c, K, iters=3)
    finally:
        del os.environ["TMVB_TILE_CAP_MAX"]
    np.testing.assert_allclose(trace, ref, rtol=ELBO_RTOL)
    np.testing.assert_allclose(model.tau, st.tau, rtol=5e-3, atol=1e-5)
    np.testing.assert_allclose(model.gamma.T, st.gamma, rtol=5e-3, atol=1e-4)
    model2, trace2, _, _, _ = _run_pair(tm, orc, c, K, iters=3)     # the same through full tiles
    np.testing.assert_allclose(trace2, trace, rtol=1e-6)


def test_flda_checkelbo_tol_and_second_call(tm, orc):
    c = tm.synth.gencorp_lda(M=120, V=400, K=5, seed=9)
    K = 7
    beta0, kappa0 = _init(tm, c, K)
    model = tm.gpufLDA(tm.Corpus.from_csr(c), K)
    model.beta = np.array(beta0.T, order="F", copy=True)
    model.kappa = kappa0.copy()
    tr = []
    tm.train(model, iter=4, tol=0.0, checkelbo=2, printelbo=False, trace=tr)
    st = orc.FLDAState(K, c.M, c.V, len(c.terms), beta0, kappa0)
    ref, _, _ = orc.flda_train(st, c.N_cumsum, c.terms, c.counts, iter=4, tol=0.0, checkelbo=2)
    np.testing.assert_allclose(tr, ref[np.isfinite(ref)], rtol=ELBO_RTOL)
    # a second train! call continues from the downloaded state (tau, eta, kappa, alpha persist)
    tr2 = []
    tm.train(model, iter=2, tol=0.0, printelbo=False, trace=tr2)
    ref2, _, _ = orc.flda_train(st, c.N_cumsum, c.terms, c.counts, iter=2, tol=0.0)
    np.testing.assert_allclose(tr2[1:], ref2[1:], rtol=ELBO_RTOL)
    # a huge tolerance stops after the first checked iteration (check_elbo!, modelutils.jl:581-583)
    m3 = tm.gpufLDA(tm.Corpus.from_csr(c), K, seed=1)
    tr3 = []
    tm.train(m3, iter=10, tol=1e12, printelbo=False, trace=tr3)
    assert len(tr3) == 2


def test_flda_argument_errors(tm):
    c = tm.synth.gencorp_lda(M=10, V=50, K=3, seed=0)
    with pytest.raises(ValueError, match="positive integer"):
        tm.gpufLDA(tm.Corpus.from_csr(c), 0)
    m = tm.gpufLDA(tm.Corpus.from_csr(c), 3)
    with pytest.raises(ValueError, match="tolerance parameters must be nonnegative"):
        tm.train(m, iter=1, tol=-1.0, printelbo=False)
    with pytest.raises(ValueError, match="checkelbo"):
        tm.train(m, iter=1, checkelbo=0, printelbo=False)
    m.eta = 1.5
    with pytest.raises(tm.TopicModelError, match="eta must belong"):
        tm.train(m, iter=1, printelbo=False)
    m.eta = 0.5
    m.kappa = np.full(c.V, 0.5, np.float32)
    with pytest.raises(tm.TopicModelError, match="kappa must be a probability vector"):
        tm.train(m, iter=1, printelbo=False)
    m = tm.gpufLDA(tm.Corpus.from_csr(c), 3)
    m.tau = np.full(len(c.terms), 1.5, np.float32)
    with pytest.raises(tm.TopicModelError, match="tau must contain probabilities"):
        tm.train(m, iter=1, printelbo=False)
    m = tm.gpufLDA(tm.Corpus.from_csr(c), 3)
    m.gamma[1, 2] = -1.0
    with pytest.raises(tm.TopicModelError, match="gamma must be positive"):
        tm.train(m, iter=1, printelbo=False)


def test_flda_nsf_size_parity(tm, orc):
    """NSF-size filtered LDA, K = 50: ELBO within 1e-4 relative of the CPU oracle (all host threads) at every iteration."""
    c = tm.synth.load_packed("nsf") or tm.synth.nsf_shaped()
    model, trace, st, ref, sweeps = _run_pair(tm, orc, c, 50, iters=3, nthreads=orc.host_threads())
    rel = np.abs(trace - ref) / np.abs(ref)
    stt = model.stats()
    print("NSF-size fLDA ELBO gpu   ", trace.tolist())
    print("NSF-size fLDA ELBO oracle", ref.tolist())
    print("rel diff", rel.tolist(), "estep_ms %.3f mstep_ms %.3f sweeps/doc %.2f eta %.5f (oracle %.5f)" % (
        stt.estep_ms, stt.mstep_ms, stt.sweeps / c.M, model.eta, st.eta[0]))
    assert np.all(rel < ELBO_RTOL)
    assert abs(model.eta - st.eta[0]) < 1e-4
    tm.check_model(model)


# ---- filtered CTM (gpufCTM, fCTM.jl) -------------------------------------------------------------------------------------------
def _run_pair_fctm(tm, orc, c, K, iters, nthreads=1):
    beta0, kappa0 = _init(tm, c, K)
    model = tm.gpufCTM(tm.Corpus.from_csr(c), K)
    model.beta = np.array(beta0.T, order="F", copy=True)
    model.kappa = kappa0.copy()
    trace = []
    tm.train(model, iter=iters, tol=0.0, checkelbo=1, printelbo=False, trace=trace)
    st = orc.FCTMState(K, c.M, c.V, len(c.terms), beta0, kappa0)
    ref, sweeps, done = orc.fctm_train(st, c.N_cumsum, c.terms, c.counts, iter=iters, tol=0.0, nthreads=nthreads)
    return model, np.array(trace), st, ref[np.isfinite(ref)], sweeps


@pytest.mark.parametrize("K", [5, 1, 3, 8, 30, 64, 100])
def test_fctm_elbo_trajectory_small(tm, orc, K):
    c = tm.synth.gencorp_lda(M=60, V=250, K=4, seed=K + 1)
    model, trace, st, ref, sweeps = _run_pair_fctm(tm, orc, c, K, iters=4)
    n = min(len(trace), len(ref))
    assert n >= 3 and (K == 1 or n == 5)
    np.testing.assert_allclose(trace[:n], ref[:n], rtol=ELBO_RTOL)
    if n < 5:
        return
    assert model.eta == 0.5                                   # update_eta! is commented out of the loop (fCTM.jl:279)
    np.testing.assert_allclose(model.kappa, st.kappa, rtol=5e-3, atol=1e-7)
    np.testing.assert_allclose(model.beta.T, st.beta, rtol=2e-2, atol=1e-5)
    np.testing.assert_allclose(model.mu, st.mu, rtol=5e-3, atol=5e-4)
    np.testing.assert_allclose(model.sigma, st.sigma, rtol=5e-3, atol=5e-4)
    np.testing.assert_allclose(model.lam.T, st.lam, rtol=5e-3, atol=5e-3)
    np.testing.assert_allclose(model.lam_old.T, st.lam_old, rtol=5e-3, atol=5e-3)
    np.testing.assert_allclose(model.vsq.T, st.vsq, rtol=5e-3, atol=1e-5)
    np.testing.assert_allclose(model.logzeta, st.logzeta, rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(model.tau, st.tau, rtol=5e-3, atol=1e-5)
    np.testing.assert_allclose(model.tau_old, st.tau_old, rtol=5e-3, atol=1e-5)
    tm.check_model(model)
    assert abs(model.stats().sweeps - int(sweeps[-1])) <= max(3, 0.02 * sweeps[-1])


def test_fctm_citeulike_size_parity(tm, orc):
    """CiteULike-size filtered CTM, K = 30: ELBO within 1e-4 relative of the CPU oracle (all host threads) at every iteration."""
    c = tm.synth.load_packed("citeu") or tm.synth.citeu_shaped()
    c = tm.synth.CSR(c.M, c.V, c.N_cumsum, c.terms, c.counts)      # the reader lists play no part
    model, trace, st, ref, sweeps = _run_pair_fctm(tm, orc, c, 30, iters=2, nthreads=orc.host_threads())
    rel = np.abs(trace - ref) / np.abs(ref)
    stt = model.stats()
    print("CiteULike-size fCTM ELBO gpu   ", trace.tolist())
    print("CiteULike-size fCTM ELBO oracle", ref.tolist())
    print("rel diff", rel.tolist(), "estep_ms %.3f sweeps/doc %.2f" % (stt.estep_ms, stt.sweeps / c.M))
    assert np.all(rel < ELBO_RTOL)
    tm.check_model(model)


def test_fctm_argument_errors(tm):
    c = tm.synth.gencorp_lda(M=10, V=50, K=3, seed=0)
    m = tm.gpufCTM(tm.Corpus.from_csr(c), 3)
    m.eta = -0.1
    with pytest.raises(tm.TopicModelError, match="eta must belong"):
        tm.train(m, iter=1, printelbo=False)
    m = tm.gpufCTM(tm.Corpus.from_csr(c), 3)
    m.tau = np.full(len(c.terms), 2.0, np.float32)
    with pytest.raises(tm.TopicModelError, match="tau must contain probabilities"):
        tm.train(m, iter=1, printelbo=False)
    with pytest.raises(ValueError, match="K <= 128"):
        tm.train(tm.gpufCTM(tm.Corpus.from_csr(c), 129), iter=1, printelbo=False)


def test_filtered_predict(tm, orc):
    """predict(corp, train_model) for the filtered models (modelutils.jl:857-884, 915-944): the inner loop on unseen documents with
    frozen globals, no scatter.  fLDA is checked against the oracle's inner loop on the same globals (one outer iteration of the
    oracle from the trained globals: its E-step is exactly predict's loop)."""
    c = tm.synth.gencorp_lda(M=120, V=300, K=4, seed=21)
    K = 6
    train_c, new_c = tm.synth.take_docs(c, np.arange(0, 90)), tm.synth.take_docs(c, np.arange(90, 120))
    beta0, kappa0 = _init(tm, c, K)
    m = tm.gpufLDA(tm.Corpus.from_csr(train_c), K)
    m.beta, m.kappa = np.array(beta0.T, order="F", copy=True), kappa0.copy()
    tm.train(m, iter=3, tol=0.0, printelbo=False)
    p = tm.predict(tm.Corpus.from_csr(new_c), m, iter=10)
    assert p.gamma.shape == (K, new_c.M) and np.all(p.gamma > 0) and np.all((p.tau >= 0) & (p.tau <= 1))
    np.testing.assert_allclose(p.beta, m.beta)                       # globals untouched: nothing was scattered, no M-step ran
    np.testing.assert_allclose(p.kappa, m.kappa)
    st = orc.FLDAState(K, new_c.M, new_c.V, len(new_c.terms), np.asarray(m.beta, np.float64).T, np.asarray(m.kappa, np.float64),
                       alpha=np.asarray(m.alpha, np.float64), eta=m.eta)
    st.tau[:] = 0.5                                                    # the new model's tau is the constructor's (fLDA.jl:50), not eta-trained
    st.tau_old[:] = 0.5
    orc.flda_train(st, new_c.N_cumsum, new_c.terms, new_c.counts, iter=1, tol=0.0, viter=10, checkelbo=float("inf"))
    # a term the training documents never contained has beta = 0 in every topic AND kappa = 0: the reference's update_tau!
    # (fLDA.jl:193) then evaluates 0 * 0^(-phi) = 0 * Inf = NaN and the document's gamma is lost (the oracle reproduces that);
    # the device works from log2(beta + eps) and stays finite.  Compare the documents the reference can handle.
    ok = np.isfinite(st.gamma).all(axis=1)
    assert ok.sum() >= new_c.M // 2 and np.all(np.isfinite(p.gamma)) and np.all(np.isfinite(p.tau))
    np.testing.assert_allclose(p.gamma.T[ok], st.gamma[ok], rtol=5e-3, atol=1e-4)
    tok_ok = np.repeat(ok, np.diff(new_c.N_cumsum))
    np.testing.assert_allclose(p.tau[tok_ok], st.tau[tok_ok], rtol=5e-3, atol=1e-5)
    d0 = int(np.flatnonzero(ok)[0])
    np.testing.assert_allclose(tm.topicdist(p, d0), st.gamma[d0] / st.gamma[d0].sum(), rtol=5e-3)
    # fCTM: runs, finite, globals untouched
    mc = tm.gpufCTM(tm.Corpus.from_csr(train_c), K)
    mc.beta, mc.kappa = np.array(beta0.T, order="F", copy=True), kappa0.copy()
    tm.train(mc, iter=2, tol=0.0, printelbo=False)
    pc = tm.predict(tm.Corpus.from_csr(new_c), mc, iter=10)
    assert np.all(np.isfinite(pc.lam)) and np.all(pc.vsq > 0) and np.all((pc.tau >= 0) & (pc.tau <= 1))
    np.testing.assert_allclose(pc.beta, mc.beta)


@pytest.mark.parametrize("case", ["flda_cfg", "fctm_cfg"])
def test_filtered_against_committed_golden(tm, case):
    """tests/golden/{flda_cfg,fctm_cfg}.npz: generated by tools/make_golden.py from both oracle restatements (C and numpy twin);
    the same inputs are exported for the real reference under tests/golden/reference/inputs (tools/reference_golden.jl)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", case + ".npz"))
    K, V = int(g["K"]), int(g["V"])
    c = tm.synth.CSR(len(g["N_cumsum"]) - 1, V, g["N_cumsum"], g["terms"].astype(np.int64), g["counts"].astype(np.int64))
    model = (tm.gpufLDA if case == "flda_cfg" else tm.gpufCTM)(tm.Corpus.from_csr(c), K)
    model.beta = np.array(g["beta0"].T, dtype=np.float32, order="F")
    model.kappa = g["kappa0"].astype(np.float32)
    tr = []
    tm.train(model, iter=len(g["elbo"]) - 1, tol=0.0, printelbo=False, trace=tr)
    np.testing.assert_allclose(tr, g["elbo"][: len(tr)], rtol=ELBO_RTOL)
    assert len(tr) == len(g["elbo"])
    np.testing.assert_allclose(model.kappa, g["kappa"], rtol=5e-3, atol=1e-7)
    np.testing.assert_allclose(model.beta.T, g["beta"], rtol=1e-2, atol=1e-6)
    np.testing.assert_allclose(model.tau, g["tau"], rtol=5e-3, atol=1e-5)
    if case == "flda_cfg":
        assert abs(model.eta - float(np.ravel(g["eta"])[0])) < 1e-4
        np.testing.assert_allclose(model.alpha, g["alpha"], rtol=2e-3)
        np.testing.assert_allclose(model.gamma.T, g["gamma"], rtol=5e-3, atol=1e-4)
    else:
        np.testing.assert_allclose(model.mu, g["mu"], rtol=5e-3, atol=1e-4)
        np.testing.assert_allclose(model.sigma, g["sigma"], rtol=1e-2, atol=1e-3)
