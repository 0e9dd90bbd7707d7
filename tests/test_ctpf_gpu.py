"""GPU parity tests for the CTPF path (CUDA through the C ABI vs the fp64 CPU oracle, same seeded inputs).
The oracle evaluates the reference's long-form ELBO (Binomial x lnGamma sums); the device uses the closed form."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ELBO_RTOL = 2e-5   # north star: 1e-4


def _run_pair(tm, orc, c, K, iters, seed=7, nthreads=1):
    alef0 = tm.synth.init_alef(K, c.V, seed=seed).astype(np.float32)    # (V, K)
    model = tm.gpuCTPF(tm.Corpus.from_csr(c), K)
    model.alef = np.array(alef0.T, order="F", copy=True)
    trace = []
    tm.train(model, iter=iters, tol=0.0, checkelbo=1, printelbo=False, trace=trace)
    st = orc.CTPFState(K, c.M, c.V, c.U, alef0)
    ref, sweeps, done = orc.ctpf_train(st, c, iter=iters, tol=0.0, nthreads=nthreads)
    return model, np.array(trace), st, ref[np.isfinite(ref)], sweeps


@pytest.mark.parametrize("K", [5, 1, 3, 8, 20, 30, 50, 100])
def test_ctpf_elbo_trajectory_small(tm, orc, K):
    c = tm.synth.gencorp_ctpf(M=60, V=300, U=40, K=4, seed=0)
    model, trace, st, ref, _ = _run_pair(tm, orc, c, K, iters=5)
    assert len(trace) == len(ref) >= 2
    np.testing.assert_allclose(trace, ref, rtol=ELBO_RTOL)
    np.testing.assert_allclose(model.alef.T, st.alef, rtol=5e-3, atol=1e-5)
    np.testing.assert_allclose(model.he.T, st.he, rtol=5e-3, atol=1e-5)
    # per-document shapes after 5 outer iterations: fp32 summation order moves single entries by up to ~6e-3 relative
    np.testing.assert_allclose(model.gimel.T, st.gimel, rtol=1e-2, atol=1e-4)
    np.testing.assert_allclose(model.zayin.T, st.zayin, rtol=1e-2, atol=1e-4)
    for n in ("bet", "vav", "dalet", "het"):
        np.testing.assert_allclose(getattr(model, n), getattr(st, n), rtol=1e-3)
        np.testing.assert_allclose(getattr(model, n + "_old"), getattr(st, n + "_old"), rtol=1e-3)
    np.testing.assert_allclose(model.gimel_old.T, st.gimel_old, rtol=5e-3, atol=1e-4)
    np.testing.assert_allclose(model.alef_old.T, st.alef_old, rtol=5e-3, atol=1e-5)
    tm.check_model(model)
    assert np.all(model.alef > 0) and np.all(model.he > 0) and np.all(model.gimel > 0) and np.all(model.zayin > 0)


def test_ctpf_fused_elbo_equals_standalone(tm):
    c = tm.synth.gencorp_ctpf(M=100, V=250, U=60, K=4, seed=3)
    K = 7
    model = tm.gpuCTPF(tm.Corpus.from_csr(c), K, seed=2)
    model.update_buffer()
    for it in range(3):
        model.estep(10, 1.0 / K**2, want_elbo=True)
        model.mstep()
        e0, e1 = model.update_elbo(0), model.update_elbo(1)
        assert abs(e0 - e1) <= 5e-6 * abs(e1), (it, e0, e1)


def test_ctpf_ratings_counts_and_ragged_lists(tm, orc):
    """ratings > 1 (the Binomial sums of CTPF.jl:116 no longer vanish), documents without readers, without terms,
    and with more readers / terms than any tile holds (overflow paths)."""
    rng = np.random.default_rng(1)
    V, U, K = 700, 300, 6
    nlen = [0, 1, 30, 600, 12, 64, 65, 5]
    rlen = [3, 0, 100, 2, 0, 17, 1, 250]
    def lists(lens, hi, vmax):
        ids, vals, off = [], [], [0]
        for L in lens:
            ids.append(rng.choice(hi, size=L, replace=False))
            vals.append(rng.integers(1, vmax, size=L))
            off.append(off[-1] + L)
        return np.array(off, np.int64), np.concatenate(ids).astype(np.int64), np.concatenate(vals).astype(np.int64)
    off, t, cn = lists(nlen, V, 6)
    roff, rd, rt = lists(rlen, U, 4)
    c = tm.synth.CSR(len(nlen), V, off, t, cn, U, roff, rd, rt)
    model, trace, st, ref, _ = _run_pair(tm, orc, c, K, iters=3)
    np.testing.assert_allclose(trace, ref, rtol=ELBO_RTOL)
    np.testing.assert_allclose(model.he.T, st.he, rtol=5e-3, atol=1e-5)
    np.testing.assert_allclose(model.zayin.T, st.zayin, rtol=5e-3, atol=1e-4)
    sc = model.scores()
    assert sc.shape == (c.M, U) and np.all(np.isfinite(sc)) and np.all(sc > 0)
    for i in range(K):
        want = np.argsort(model.alef[i, :], kind="stable")[::-1] + 1
        np.testing.assert_array_equal(np.asarray(model.topics[i]), want)


def test_ctpf_argument_errors(tm):
    c = tm.synth.gencorp_ctpf(M=10, V=50, U=8, K=3, seed=0)
    with pytest.raises(ValueError):
        tm.gpuCTPF(tm.Corpus.from_csr(c), 0)
    m = tm.gpuCTPF(tm.Corpus.from_csr(c), 3)
    with pytest.raises(ValueError):
        tm.train(m, vtol=-1.0)
    m.a = -0.1
    with pytest.raises(tm.TopicModelError, match="a must be positive"):
        tm.train(m, iter=1, printelbo=False)
    m = tm.gpuCTPF(tm.Corpus.from_csr(c), 3)
    m.alef[0, 0] = 0.0
    with pytest.raises(tm.TopicModelError, match="positive"):
        tm.train(m, iter=1, printelbo=False)
    m = tm.gpuCTPF(tm.Corpus.from_csr(c), 3)
    m.vav = np.array([1.0, np.inf, 1.0], dtype=np.float32)
    with pytest.raises(tm.TopicModelError, match="vav must be finite"):
        tm.train(m, iter=1, printelbo=False)


def test_ctpf_citeulike_size_parity(tm, orc):
    """BASELINE config 3: gpuCTPF K=30 on CiteULike with 5551 users (packed real corpus when present, else
    CiteULike-shaped synthetic).  ELBO within 1e-4 relative of the CPU oracle at every iteration (asserted 2e-5)."""
    c = tm.synth.load_packed("citeu") or tm.synth.citeu_shaped()
    model, trace, st, ref, sweeps = _run_pair(tm, orc, c, 30, iters=3, nthreads=orc.host_threads())
    rel = np.abs(trace - ref) / np.abs(ref)
    print("CiteULike-size CTPF ELBO gpu   ", trace.tolist())
    print("CiteULike-size CTPF ELBO oracle", ref.tolist())
    print("rel diff", rel.tolist(), "estep_ms", model.stats().estep_ms)
    assert np.all(rel < ELBO_RTOL)
    tm.check_model(model)


def test_ctpf_against_committed_golden(tm):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ctpf_cfg.npz"))
    K, V, U = int(g["K"]), int(g["V"]), int(g["U"])
    c = tm.synth.CSR(len(g["N_cumsum"]) - 1, V, g["N_cumsum"], g["terms"].astype(np.int64), g["counts"].astype(np.int64), U,
                     g["R_cumsum"], g["readers"].astype(np.int64), g["ratings"].astype(np.int64))
    model = tm.gpuCTPF(tm.Corpus.from_csr(c), K)
    model.alef = np.array(g["alef0"].T, dtype=np.float32, order="F")
    tr = []
    tm.train(model, iter=len(g["elbo"]) - 1, tol=0.0, printelbo=False, trace=tr)
    np.testing.assert_allclose(tr, g["elbo"][: len(tr)], rtol=ELBO_RTOL)
    np.testing.assert_allclose(model.vav, g["vav"], rtol=2e-3)


# ---- the recommendation step of train!(::gpuCTPF) (gpuCTPF.jl:709-731) on the device: tmvb_ctpf_recs ----------------------
def _rank_from(scores_col, excluded0, n):
    keep = np.ones(n, bool)
    keep[excluded0] = False
    idx = np.flatnonzero(keep)
    return idx[np.argsort(scores_col[idx], kind="stable")[::-1]] + 1


def _check_rankings_against(sc, ur, dr, c):
    """bit-exact: the rankings are findall(.)[reverse(sortperm(.))] of the score matrix the device itself returned"""
    Rc, readers = np.asarray(c.R_cumsum), np.asarray(c.readers)
    libs = [[] for _ in range(c.U)]
    for d in range(c.M):
        for u in readers[Rc[d]:Rc[d + 1]]:
            libs[u].append(d)
    for u in range(c.U):
        np.testing.assert_array_equal(ur[u], _rank_from(sc[:, u], libs[u], c.M))
    for d in range(c.M):
        np.testing.assert_array_equal(dr[d], _rank_from(sc[d, :], readers[Rc[d]:Rc[d + 1]], c.U))


@pytest.mark.parametrize("M,U,K", [(60, 40, 4), (300, 130, 30), (129, 257, 33), (5, 3, 2), (700, 515, 100)])
def test_ctpf_recs_small(tm, orc, M, U, K):
    """scores: tensor-core contraction (tf32 hi/lo split) == CUDA-core fp32 contraction == fp64 oracle to fp32 rounding;
    urecs / drecs: exactly the reference's ranking of the returned scores (ties included), and the oracle's ranking wherever
    the oracle's scores are not tied to within fp32 rounding."""
    c = tm.synth.gencorp_ctpf(M=M, V=200, U=U, K=3, seed=M + U)
    model, trace, st, ref, _ = _run_pair(tm, orc, c, K, iters=2)
    sc, ur, dr = model.update_recs()
    sc1, ur1, dr1 = model.update_recs(mode=1)
    assert sc.shape == (c.M, c.U) and sc.flags.f_contiguous
    np.testing.assert_allclose(sc, sc1, rtol=1e-5, atol=0)     # K <= 100 positive terms: fp32 summation order only
    st32 = orc.CTPFState(K, c.M, c.V, c.U, st.alef)
    for n in ("he", "vav", "gimel", "zayin", "dalet", "het"):
        setattr(st32, n, np.asarray(getattr(model, n), dtype=np.float64).T if getattr(model, n).ndim == 2 else np.asarray(getattr(model, n), np.float64))
    so, uo, do = orc.ctpf_recs(st32, c)                     # fp64 loops over the device's own (downloaded) state
    np.testing.assert_allclose(sc, so, rtol=1e-5)
    np.testing.assert_allclose(sc1, so, rtol=1e-5)
    print("scores vs fp64: tensor cores %.2e, CUDA cores %.2e (max rel)" % (np.max(np.abs(sc - so) / so), np.max(np.abs(sc1 - so) / so)))
    _check_rankings_against(sc, ur, dr, c)
    _check_rankings_against(sc1, ur1, dr1, c)
    nz = c.M * c.U - len(c.readers)
    assert sum(len(x) for x in ur) == nz == sum(len(x) for x in dr)
    for u in range(c.U):                                    # against the oracle's rankings: swaps only between near-ties
        bad = np.flatnonzero(ur[u] != uo[u])
        if len(bad):
            a, b = so[ur[u][bad] - 1, u], so[uo[u][bad] - 1, u]
            np.testing.assert_allclose(a, b, rtol=1e-5)
    # the model after the oracle's own training gives the same scores to the ELBO-level tolerance
    s_ref, _, _ = orc.ctpf_recs(st, c)
    np.testing.assert_allclose(sc, s_ref, rtol=2e-2, atol=1e-6)


def test_ctpf_recs_ties_and_partial_outputs(tm):
    """small-integer state: many exactly tied scores (exact in tf32 hi/lo and fp32), so the rankings must equal the
    reference's reversed stable sort element for element; every output is optional."""
    c = tm.synth.gencorp_ctpf(M=90, V=80, U=70, K=3, seed=5)
    K = 6
    m = tm.gpuCTPF(tm.Corpus.from_csr(c), K, seed=1)
    rng = np.random.default_rng(0)
    for n, cols in (("he", c.U), ("gimel", c.M), ("zayin", c.M)):
        setattr(m, n, np.asfortranarray(rng.integers(1, 4, size=(K, cols)).astype(np.float32)))
    for n in ("vav", "dalet", "het"):
        setattr(m, n, rng.integers(1, 3, size=K).astype(np.float32))
    want = m.scores()
    sc, ur, dr = m.update_recs()
    np.testing.assert_array_equal(sc, want)
    _check_rankings_against(sc, ur, dr, c)
    for u in range(c.U):
        np.testing.assert_array_equal(ur[u], m.urecs[u])
    for d in range(c.M):
        np.testing.assert_array_equal(dr[d], m.drecs[d])
    s2, u2, d2 = m.update_recs(scores=False, urecs=True, drecs=False)
    assert s2 is None and d2 is None
    for u in range(c.U):
        np.testing.assert_array_equal(u2[u], ur[u])
    s3, u3, d3 = m.update_recs(scores=False, urecs=False, drecs=True)
    for d in range(c.M):
        np.testing.assert_array_equal(d3[d], dr[d])


def test_ctpf_recs_citeulike_size(tm):
    """BASELINE config 3 shape (16 980 x 5 551, K = 30): tensor-core scores == CUDA-core scores, rankings of sampled
    users / documents == the reference's loops over the returned scores; the end of train!(recs=True)."""
    import time
    c = tm.synth.load_packed("citeu") or tm.synth.citeu_shaped()
    model = tm.gpuCTPF(tm.Corpus.from_csr(c), 30, seed=3)
    tm.train(model, iter=2, tol=0.0, printelbo=False, recs=True)
    t0 = time.perf_counter()
    sc, ur, dr = model.update_recs()
    t1 = time.perf_counter()
    sc1, _, _ = model.update_recs(urecs=False, drecs=False, mode=1)
    print("update_recs (scores + urecs + drecs, M=%d U=%d): %.1f ms" % (c.M, c.U, (t1 - t0) * 1e3))
    np.testing.assert_allclose(sc, sc1, rtol=1e-5)
    np.testing.assert_array_equal(sc, model.scores_)
    Eeta = (model.he / model.vav[:, None]).astype(np.float64)
    X = (model.gimel / model.dalet[:, None] + model.zayin / model.het[:, None]).astype(np.float64)
    rng = np.random.default_rng(0)
    Rc, readers = np.asarray(c.R_cumsum), np.asarray(c.readers)
    for d in rng.choice(c.M, 40, replace=False):
        np.testing.assert_allclose(sc[d, :], X[:, d] @ Eeta, rtol=1e-5)
        np.testing.assert_array_equal(dr[d], _rank_from(sc[d, :], readers[Rc[d]:Rc[d + 1]], c.U))
    for u in rng.choice(c.U, 40, replace=False):
        np.testing.assert_array_equal(ur[u], _rank_from(sc[:, u], np.asarray(model.libs[u], dtype=np.int64) - 1, c.M))
    assert sum(len(x) for x in ur) == c.M * c.U - len(readers) == sum(len(x) for x in dr)
