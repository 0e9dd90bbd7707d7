"""GPU parity tests for the LDA path: CUDA (through the C ABI) vs the fp64 CPU oracle on the same
seeded inputs.  Tolerances: ELBO 1e-4 relative at every outer iteration is the north-star bar; the
engine is fp32 per element with fp64 accumulation and lands around 1e-7, so the tests assert 2e-6."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ELBO_RTOL = 2e-6      # asserted; north star allows 1e-4


def _run_pair(tm, orc, c, K, iters, seed=7, viter=10, checkelbo=1, nthreads=1, alpha0=None):
    beta0 = tm.synth.init_beta(K, c.V, seed=seed).astype(np.float32)   # (V, K)
    model = tm.gpuLDA(tm.Corpus.from_csr(c), K)
    model.beta = np.array(beta0.T, order="F", copy=True)
    if alpha0 is not None:
        model.alpha = np.asarray(alpha0, dtype=np.float32)
    trace = []
    tm.train(model, iter=iters, tol=0.0, viter=viter, checkelbo=checkelbo, printelbo=False, trace=trace)
    st = orc.LDAState(K, c.M, c.V, beta=beta0, alpha=None if alpha0 is None else np.asarray(alpha0, dtype=np.float32))
    ref, sweeps, done = orc.lda_train(st, c.N_cumsum, c.terms, c.counts, iter=iters, tol=0.0, viter=viter,
                                      checkelbo=checkelbo, nthreads=nthreads)
    return model, np.array(trace), st, ref[np.isfinite(ref)], sweeps


@pytest.mark.parametrize("K", [5, 1, 2, 8, 9, 30, 50, 64, 100])
def test_elbo_trajectory_small(tm, orc, K):
    c = tm.synth.gencorp_lda(M=100, V=500, K=5, seed=0)
    model, trace, st, ref, _ = _run_pair(tm, orc, c, K, iters=8)
    if K == 1:   # the ELBO is flat after one iteration: fp noise decides when delta < tol=0 stops either side
        n = min(len(trace), len(ref))
        assert n >= 2
        np.testing.assert_allclose(trace[:n], ref[:n], rtol=ELBO_RTOL)
        return
    assert len(trace) == len(ref) == 9
    np.testing.assert_allclose(trace, ref, rtol=ELBO_RTOL)
    np.testing.assert_allclose(model.alpha, st.alpha, rtol=2e-4)
    np.testing.assert_allclose(model.beta.T, st.beta, rtol=5e-3, atol=1e-9)
    np.testing.assert_allclose(model.gamma.T, st.gamma, rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(model.Elogtheta.T, st.Elogtheta, rtol=2e-3, atol=2e-4)
    # invariants of check_model(::gpuLDA) (modelutils.jl:255-279)
    tm.check_model(model)
    assert np.all(model.Elogtheta <= 0) and np.all(model.gamma > 0)


def test_sweep_counts_match(tm, orc):
    c = tm.synth.gencorp_lda(M=300, V=800, K=8, seed=3)
    K = 8
    beta0 = tm.synth.init_beta(K, c.V).astype(np.float32)
    model = tm.gpuLDA(tm.Corpus.from_csr(c), K)
    model.beta = np.array(beta0.T, order="F", copy=True)
    model.update_buffer()
    st = orc.LDAState(K, c.M, c.V, beta=beta0)
    for it in range(4):
        model.estep(10, 1.0 / K**2, want_elbo=False)
        model.update_beta()
        model.update_alpha(1000, 1.0 / K**2)
        gs = model.stats().sweeps
        _, sw, _ = orc.lda_train(st, c.N_cumsum, c.terms, c.counts, iter=1, tol=0.0, checkelbo=math.inf)
        # per-document early exit follows the same rule; fp32 rounding may flip a handful of documents
        assert abs(gs - int(sw[0])) <= max(3, 0.01 * sw[0]), (it, gs, sw)


def test_fused_elbo_equals_standalone(tm, orc):
    """mode 0 (partials fused into the E-step/M-step) == mode 1 (one table-assisted pass over the device state)
    == mode 2 (update_elbo!, LDA.jl:50-93, restated literally in fp64 on the device)."""
    c = tm.synth.gencorp_lda(M=200, V=600, K=6, seed=5)
    K = 10
    model = tm.gpuLDA(tm.Corpus.from_csr(c), K, seed=11)
    model.update_buffer()
    for it in range(3):
        model.estep(10, 1.0 / K**2, want_elbo=True)
        model.update_beta()
        model.update_alpha(1000, 1.0 / K**2)
        e0 = model.update_elbo(0)
        e1 = model.update_elbo(1)
        e2 = model.update_elbo(2)
        assert abs(e0 - e2) <= 1e-6 * abs(e2), (it, e0, e2)
        assert abs(e1 - e2) <= 1e-6 * abs(e2), (it, e1, e2)


@pytest.mark.parametrize("K", [1, 3, 10, 50, 100, 200])
def test_fresh_state_elbo_equals_literal(tm, monkeypatch, K):
    """mode 1 right after update_buffer! (beta_old == beta, Elogtheta_old == Elogtheta on the device: the `update_elbo!` that opens
    every train! call, gpuLDA.jl:353) takes the one-dot-product-per-token kernel; it must equal the table-assisted pass
    (TMVB_ELBO_NOFRESH=1) and the literal fp64 restatement (mode 2) on the same re-uploaded state."""
    c = tm.synth.gencorp_lda(M=250, V=600, K=6, seed=K)
    model = tm.gpuLDA(tm.Corpus.from_csr(c), K, seed=3)
    tm.train(model, iter=2, tol=0.0, printelbo=False)        # a non-trivial state on the host (Elogtheta, gamma, beta, alpha)
    model.update_buffer()                                     # re-upload: the lagged copies now equal the current ones
    e_fresh = model.update_elbo(1)
    monkeypatch.setenv("TMVB_ELBO_NOFRESH", "1")
    e_table = model.update_elbo(1)
    monkeypatch.delenv("TMVB_ELBO_NOFRESH")
    e_lit = model.update_elbo(2)
    assert abs(e_fresh - e_lit) <= 1e-6 * abs(e_lit), (e_fresh, e_lit)
    assert abs(e_table - e_lit) <= 1e-6 * abs(e_lit), (e_table, e_lit)
    # after a step the copies differ again and mode 1 falls back to the table pass
    model.estep(10, 1.0 / K**2, want_elbo=True)
    model.update_beta()
    model.update_alpha(1000, 1.0 / K**2)
    assert abs(model.update_elbo(1) - model.update_elbo(2)) <= 1e-6 * abs(model.update_elbo(2))


def test_fused_iteration_equals_separate_steps(tm, monkeypatch):
    """tmvb_lda_iterate (one CUDA graph launch per outer iteration: what train() uses) == estep + mstep + update_alpha + elbo."""
    c = tm.synth.gencorp_lda(M=300, V=700, K=6, seed=21)
    K = 12

    def run(unfused):
        monkeypatch.setenv("TMVB_UNFUSED", "1" if unfused else "0")
        model = tm.gpuLDA(tm.Corpus.from_csr(c), K, seed=5)
        tr = []
        tm.train(model, iter=5, tol=0.0, printelbo=False, trace=tr, checkelbo=1)
        assert model.can_iterate() == (not unfused)
        return np.array(tr), model.alpha.copy(), model.beta.copy(), model.gamma.copy()

    e0, a0, b0, g0 = run(True)
    e1, a1, b1, g1 = run(False)
    np.testing.assert_allclose(e1, e0, rtol=1e-6)          # same kernels, same order: only the atomics' order differs (2e-7 seen after 5 iterations)
    np.testing.assert_allclose(a1, a0, rtol=1e-5)
    np.testing.assert_allclose(b1, b0, rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(g1, g0, rtol=1e-4, atol=1e-6)
    # checkelbo = 2: iterations without an ELBO do not synchronise, the traced ones agree
    monkeypatch.setenv("TMVB_UNFUSED", "0")
    model = tm.gpuLDA(tm.Corpus.from_csr(c), K, seed=5)
    tr = []
    tm.train(model, iter=4, tol=0.0, printelbo=False, trace=tr, checkelbo=2)
    np.testing.assert_allclose(tr[1:], e0[[2, 4]], rtol=1e-6)


def test_int64_and_int32_corpus_entry_points_agree(tm, monkeypatch):
    """tmvb_lda_set_corpus (Int64 vectors, what update_buffer! builds) and tmvb_lda_set_corpus32 (the host mirror's packed
    cache) lay out the same device corpus: the same trajectories."""
    c = tm.synth.gencorp_lda(M=120, V=500, K=5, seed=21)
    out = []
    for force64 in ("1", "0"):
        monkeypatch.setenv("TMVB_CORPUS64", force64)
        model = tm.gpuLDA(tm.Corpus.from_csr(c), 7, seed=4)
        tr = []
        tm.train(model, iter=3, tol=0.0, printelbo=False, trace=tr)
        out.append((tr, model.gamma.copy(), model.beta.copy()))
    # equal up to the order of the fp32 atomic reductions into the statistics
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=1e-6)
    np.testing.assert_allclose(out[0][1], out[1][1], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(out[0][2], out[1][2], rtol=1e-3, atol=1e-8)


def test_ragged_and_edge_documents(tm, orc):
    """empty documents, single-token documents, a document longer than any shared-memory tile
    (overflow path), duplicate-free long counts."""
    rng = np.random.default_rng(0)
    V, K = 3000, 12
    lens = [0, 1, 1, 2, 17, 64, 65, 300, 2500, 0, 33]
    terms, counts, off = [], [], [0]
    for L in lens:
        terms.append(rng.choice(V, size=L, replace=False))
        counts.append(rng.integers(1, 46, size=L))
        off.append(off[-1] + L)
    c = tm.synth.CSR(len(lens), V, np.array(off, np.int64), np.concatenate(terms).astype(np.int64),
                     np.concatenate(counts).astype(np.int64))
    model, trace, st, ref, _ = _run_pair(tm, orc, c, K, iters=4)
    np.testing.assert_allclose(trace, ref, rtol=ELBO_RTOL)
    np.testing.assert_allclose(model.gamma.T, st.gamma, rtol=2e-3, atol=1e-6)
    # an empty document keeps gamma = alpha (+EPS) -- LDA.jl:143-146 with phi*counts = 0
    np.testing.assert_allclose(model.gamma[:, 0], st.gamma[0], rtol=1e-5)


def test_phi_materialisation_and_download_old(tm, orc):
    c = tm.synth.gencorp_lda(M=60, V=300, K=4, seed=2)
    K = 7
    model, trace, st, ref, _ = _run_pair(tm, orc, c, K, iters=3)
    phi_ref = orc.lda_phi(K, c.M, c.N_cumsum, c.terms, st.beta_old, st.Elogtheta_old)  # (nnz, K)
    phi = model.phi
    assert len(phi) == c.M
    got = np.concatenate([p.T for p in phi], axis=0)
    np.testing.assert_allclose(got.sum(axis=1), 1.0, rtol=1e-5)          # left-stochastic (modelutils.jl:274-276)
    np.testing.assert_allclose(got, phi_ref, rtol=5e-3, atol=1e-7)
    np.testing.assert_allclose(model.beta_old.T, st.beta_old, rtol=5e-3, atol=1e-9)
    np.testing.assert_allclose(model.Elogtheta_old.T, st.Elogtheta_old, rtol=2e-3, atol=2e-4)


def test_checkelbo_every_other_and_tol_stop(tm, orc):
    c = tm.synth.gencorp_lda(M=100, V=500, K=5, seed=0)
    model, trace, st, ref, _ = _run_pair(tm, orc, c, 5, iters=6, checkelbo=2)
    assert len(trace) == len(ref) == 4
    np.testing.assert_allclose(trace, ref, rtol=ELBO_RTOL)
    # tol stop: delta_elbo < tol terminates (modelutils.jl:580)
    model = tm.gpuLDA(tm.Corpus.from_csr(c), 5, seed=1)
    tr = []
    tm.train(model, iter=50, tol=1e9, printelbo=False, trace=tr)
    assert len(tr) == 2


def test_argument_errors(tm):
    c = tm.synth.gencorp_lda(M=10, V=50, K=3, seed=0)
    with pytest.raises(ValueError):
        tm.gpuLDA(tm.Corpus.from_csr(c), 0)
    m = tm.gpuLDA(tm.Corpus.from_csr(c), 3)
    with pytest.raises(ValueError):
        tm.train(m, tol=-1.0)
    with pytest.raises(ValueError):
        tm.train(m, iter=-1)
    with pytest.raises(ValueError):
        tm.train(m, checkelbo=0)
    m.alpha = np.array([1.0, -1.0, 1.0], dtype=np.float32)
    with pytest.raises(tm.TopicModelError):
        tm.train(m, iter=1)
    # element-wise invariants are checked on the device copy (modelutils.jl:264-273)
    m3 = tm.gpuLDA(tm.Corpus.from_csr(c), 3, seed=0)
    m3.gamma[1, 2] = 0.0
    with pytest.raises(tm.TopicModelError, match="gamma must be positive"):
        tm.train(m3, iter=1, printelbo=False)
    m3 = tm.gpuLDA(tm.Corpus.from_csr(c), 3, seed=0)
    m3.Elogtheta[0, 0] = np.nan
    with pytest.raises(tm.TopicModelError, match="Elogtheta must be finite"):
        tm.train(m3, iter=1, printelbo=False)
    # isstochastic(beta, dims=2) (modelutils.jl:268): row sums on the device copy
    m3 = tm.gpuLDA(tm.Corpus.from_csr(c), 3, seed=0)
    m3.beta[1, :] *= 1.01
    with pytest.raises(tm.TopicModelError, match="right stochastic"):
        tm.train(m3, iter=1, printelbo=False)
    m3 = tm.gpuLDA(tm.Corpus.from_csr(c), 3, seed=0)
    m3.beta[2, 4] = -0.25
    with pytest.raises(tm.TopicModelError, match="right stochastic"):
        tm.train(m3, iter=1, printelbo=False)
    m4 = tm.gpuCTM(tm.Corpus.from_csr(c), 3, seed=0)
    m4.beta[0, :] *= 0.9
    with pytest.raises(tm.TopicModelError, match="right stochastic"):
        tm.train(m4, iter=1, printelbo=False)
    # out-of-range term id is rejected by the library
    bad = tm.synth.CSR(1, 5, np.array([0, 2], np.int64), np.array([1, 7], np.int64), np.array([1, 1], np.int64))
    m2 = tm.gpuLDA(tm.Corpus.from_csr(bad), 2)
    with pytest.raises(ValueError):
        m2.update_buffer()


def test_nsf_shaped_full_size_parity(tm, orc):
    """BASELINE config 1 at full size: K=50 on an NSF-shaped corpus (128 804 docs x 25 319 vocab;
    the packed real NSF corpus when data/_packed/nsf.npz travelled with the snapshot).  ELBO within
    1e-4 relative of the CPU oracle at every outer iteration (north star), asserted at 2e-6."""
    c = tm.synth.load_packed("nsf") or tm.synth.nsf_shaped()
    iters = 10                                      # BASELINE configs[1]: iter = 10, checkelbo = 1
    model, trace, st, ref, sweeps = _run_pair(tm, orc, c, 50, iters=iters, nthreads=orc.host_threads())
    assert len(trace) == len(ref) == iters + 1
    rel = np.abs(trace - ref) / np.abs(ref)
    print("NSF-size ELBO gpu   ", trace.tolist())
    print("NSF-size ELBO oracle", ref.tolist())
    print("rel diff", rel.tolist())
    assert np.all(rel < ELBO_RTOL)
    assert np.all(np.diff(trace[1:]) > 0)          # CAVI ascent after the first iteration
    tm.check_model(model)


def test_topics_ranked_on_device(tm):
    """model.topics == [reverse(sortperm(beta[i,:]))] (gpuLDA.jl:374), 1-based, ties resolved like Julia's stable sort."""
    c = tm.synth.gencorp_lda(M=80, V=400, K=4, seed=9)
    K = 6
    model = tm.gpuLDA(tm.Corpus.from_csr(c), K, seed=3)
    tm.train(model, iter=3, tol=0.0, printelbo=False)
    assert len(model.topics) == K
    for i in range(K):
        want = np.argsort(model.beta[i, :], kind="stable")[::-1] + 1
        np.testing.assert_array_equal(np.asarray(model.topics[i]), want)


def test_lda_against_committed_golden(tm):
    """tests/golden/lda_cfg0.npz (generated by tools/make_golden.py from both oracle restatements)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lda_cfg0.npz"))
    K, V = int(g["K"]), int(g["V"])
    c = tm.synth.CSR(len(g["N_cumsum"]) - 1, V, g["N_cumsum"], g["terms"].astype(np.int64), g["counts"].astype(np.int64))
    model = tm.gpuLDA(tm.Corpus.from_csr(c), K)
    model.beta = np.array(g["beta0"].T, dtype=np.float32, order="F")
    tr = []
    tm.train(model, iter=20, tol=0.0, printelbo=False, trace=tr)
    np.testing.assert_allclose(tr, g["elbo"], rtol=ELBO_RTOL)
    np.testing.assert_allclose(model.alpha, g["alpha"], rtol=5e-4)
    # 20 outer iterations amplify fp32 rounding (and per-document stopping decisions taken at the vtol threshold):
    # nearly every entry agrees to 1e-2 relative, all agree to 1e-4 absolute (entries are O(1e-2))
    close = np.isclose(model.beta.T, g["beta"], rtol=1e-2, atol=1e-8)
    assert close.mean() > 0.995, close.mean()
    np.testing.assert_allclose(model.beta.T, g["beta"], rtol=1e-2, atol=1e-4)


def test_beta_parity_without_threshold_decisions(tm, orc):
    """The committed-golden test above allows a few beta entries 1e-2 relative off after 20 iterations and attributes them to
    per-document stopping decisions taken at the vtol threshold (a document whose ||dElogtheta|| lands within fp32 rounding of
    vtol makes one sweep more or fewer than in fp64).  Demonstration: the SAME run with vtol = 0 (`norm < 0` never holds:
    every document makes exactly viter sweeps on both sides) agrees entry for entry to 2e-3."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lda_cfg0.npz"))
    K, V = int(g["K"]), int(g["V"])
    c = tm.synth.CSR(len(g["N_cumsum"]) - 1, V, g["N_cumsum"], g["terms"].astype(np.int64), g["counts"].astype(np.int64))
    beta0 = g["beta0"].astype(np.float32)
    model = tm.gpuLDA(tm.Corpus.from_csr(c), K)
    model.beta = np.array(beta0.T, dtype=np.float32, order="F")
    tr = []
    tm.train(model, iter=20, tol=0.0, vtol=0.0, printelbo=False, trace=tr)
    st = orc.LDAState(K, c.M, V, beta=beta0)
    ref, sweeps, _ = orc.lda_train(st, c.N_cumsum, c.terms, c.counts, iter=20, tol=0.0, vtol=0.0)
    assert np.all(sweeps == 10 * c.M) and model.stats().sweeps == 10 * c.M
    np.testing.assert_allclose(tr, ref, rtol=ELBO_RTOL)
    np.testing.assert_allclose(model.beta.T, st.beta, rtol=2e-3, atol=1e-9)
    np.testing.assert_allclose(model.gamma.T, st.gamma, rtol=1e-3, atol=1e-6)


def test_lda_k200_layout(tm, orc):
    """K=200 (BASELINE config 5's topic count): the LPT=8 x CPL=7 lane layout and 7 topics per lane in the K phase."""
    c = tm.synth.gencorp_lda(M=120, V=900, K=12, seed=4, mean_len=80)
    model, trace, st, ref, _ = _run_pair(tm, orc, c, 200, iters=3)
    np.testing.assert_allclose(trace, ref, rtol=ELBO_RTOL)
    np.testing.assert_allclose(model.gamma.T, st.gamma, rtol=2e-3, atol=1e-6)


@pytest.mark.parametrize("K,M,lens", [(50, 3000, (1, 300)), (51, 700, (1, 120)), (5, 600, (1, 200)), (200, 400, (1, 260)), (7, 300, (0, 40))])
def test_host_mirror_equals_download(tm, monkeypatch, K, M, lens):
    """update_host! folded into the last E-step (tmvb_lda_arm_host_mirror): gamma / Elogtheta written by the E-step kernels into the
    page-locked arrays must be bit-identical to what tmvb_lda_download copies (same kernels, same values), for the hybrid kernel
    (K=50, odd K=51, K=200), the generic one (K=5, 7), documents of every length class incl. empty ones, and across repeated
    train! calls on one model (whose Elogtheta / gamma then ARE the mirrored arrays)."""
    rng = np.random.default_rng(K)
    V = 900
    n = rng.integers(lens[0], lens[1] + 1, size=M)
    n[:3] = lens[1]
    off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    terms = np.concatenate([rng.choice(V, size=k, replace=False) for k in n] + [np.zeros(0, np.int64)]).astype(np.int64)
    counts = rng.integers(1, 5, size=len(terms)).astype(np.int64)
    c = tm.synth.CSR(M, V, off, terms, counts)
    beta0 = tm.synth.init_beta(K, V, seed=3).astype(np.float32)
    lib = tm._lib.load()

    def device_copy(model):
        E, g = np.empty((K, M), np.float32, order="F"), np.empty((K, M), np.float32, order="F")
        tm._lib.check(lib.tmvb_lda_download(model._handle(), None, None, E.ctypes.data, g.ctypes.data))   # other pointers: a plain copy
        return E, g

    out = {}
    for mirror in ("1", "0"):
        monkeypatch.setenv("TMVB_HOST_MIRROR", mirror)
        model = tm.gpuLDA(tm.Corpus.from_csr(c), K)
        model.beta = np.array(beta0.T, order="F", copy=True)
        tr = []
        d2h = []
        for kw in (dict(iter=3, tol=0.0), dict(iter=2, tol=0.0),    # the second call uploads out of the mirrored arrays
                   dict(iter=5, tol=1e30)):                          # stops at k = 1 < iter: the mirror is never armed
            d0 = model.stats().d2h_bytes
            tm.train(model, printelbo=False, trace=tr, **kw)
            d2h.append(model.stats().d2h_bytes - d0)
            E, g = device_copy(model)
            # the rows the kernels wrote over the bus ARE the device rows (the statistics' atomics make two runs differ in the last
            # bits, so the bit-exact comparison is within one run)
            np.testing.assert_array_equal(model.Elogtheta, E)
            np.testing.assert_array_equal(model.gamma, g)
            tm.check_model(model)
        out[mirror] = (np.array(model.gamma), np.array(model.Elogtheta), np.array(model.beta), np.array(tr), d2h)
        model.close()
    for a, b in zip(out["1"][:4], out["0"][:4]):
        np.testing.assert_allclose(a, b, rtol=2e-4, atol=1e-6)
    assert out["1"][4] == out["0"][4]            # the bytes the kernels wrote over the bus are counted like the copies they replace


def test_host_mirror_argument_errors(tm):
    import ctypes as C
    c = tm.synth.gencorp_lda(M=50, V=200, K=4, seed=1)
    model = tm.gpuLDA(tm.Corpus.from_csr(c), 6)
    model.update_buffer()
    lib, h = tm._lib.load(), model._handle()
    pageable = np.zeros((6, 50), np.float32, order="F")
    pinned = tm._lib.pinned_empty((6, 50), np.float32, order="F")
    assert lib.tmvb_lda_arm_host_mirror(h, pageable.ctypes.data, pinned.ctypes.data) != 0
    assert "page-locked" in lib.tmvb_last_error().decode()
    assert lib.tmvb_lda_arm_host_mirror(h, pinned.ctypes.data, None) != 0
    assert lib.tmvb_lda_arm_host_mirror(h, None, None) == 0
    g = tm._lib.pinned_empty((6, 50), np.float32, order="F")
    assert lib.tmvb_lda_arm_host_mirror(h, pinned.ctypes.data, g.ctypes.data) == 0
    model.estep(10, 1e-3, want_elbo=False)
    E2, g2 = np.empty((6, 50), np.float32, order="F"), np.empty((6, 50), np.float32, order="F")
    tm._lib.check(lib.tmvb_lda_download(h, None, None, E2.ctypes.data, g2.ctypes.data))   # other pointers: a plain copy
    tm._lib.check(lib.tmvb_lda_download(h, None, None, pinned.ctypes.data, g.ctypes.data))  # the mirrored ones: only a wait
    np.testing.assert_array_equal(E2, pinned)
    np.testing.assert_array_equal(g2, g)
    assert np.all(g > 0) and np.all(pinned <= 0)
