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
