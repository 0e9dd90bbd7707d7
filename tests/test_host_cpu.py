"""CPU tests of the host mirror: Corpus/Document contract, the update_buffer! flattening, sharding,
synthetic corpora, train! argument validation, and the 2-rank (gloo) sufficient-statistic exchange."""
import os
import socket

import numpy as np
import pytest


def test_document_and_corpus_checks(tm):
    d = tm.Document(terms=[3, 1, 2], counts=[1, 2, 1], readers=[2], ratings=[5])
    assert len(d) == 3
    with pytest.raises(tm.DocumentError):
        tm.Document(terms=[0, 1])                        # keys are 1-based (Corpus.jl:42)
    with pytest.raises(tm.DocumentError):
        tm.Document(terms=[1, 2], counts=[1])
    with pytest.raises(tm.DocumentError):
        tm.Document(terms=[1], counts=[0])
    with pytest.raises(tm.CorpusError):
        tm.Corpus([tm.Document(terms=[5])], vocab=3)     # term key outside the vocabulary (Corpus.jl:117)
    corp = tm.Corpus([d, tm.Document(terms=[]), tm.Document(terms=[2, 2 + 1])], vocab=4, users=2)
    f = corp.flat()                                      # modelutils.jl:371-380
    assert f.M == 3 and f.V == 4 and f.U == 2
    np.testing.assert_array_equal(f.N_cumsum, [0, 3, 3, 5])
    np.testing.assert_array_equal(f.terms, [2, 0, 1, 1, 2])   # 0-based
    np.testing.assert_array_equal(f.counts, [1, 2, 1, 1, 1])
    np.testing.assert_array_equal(f.R_cumsum, [0, 1, 1, 1])
    np.testing.assert_array_equal(f.readers, [1])


def test_shard_is_a_partition(tm):
    c = tm.synth.gencorp_lda(M=37, V=90, K=3, seed=1)
    for world in (2, 3, 8):
        shards = [c.shard(r, world) for r in range(world)]
        assert sum(s.M for s in shards) == c.M and sum(s.nnz for s in shards) == c.nnz
        for r, s in enumerate(shards):
            docs = np.arange(r, c.M, world)
            np.testing.assert_array_equal(np.diff(s.N_cumsum), np.diff(c.N_cumsum)[docs])
            d0 = docs[0]
            np.testing.assert_array_equal(s.terms[: s.N_cumsum[1]], c.terms[c.N_cumsum[d0]: c.N_cumsum[d0 + 1]])


def test_synthetic_corpora_are_deterministic_and_condensed(tm):
    a, b = tm.synth.nsf_shaped(M=3000, V=2000), tm.synth.nsf_shaped(M=3000, V=2000)
    np.testing.assert_array_equal(a.terms, b.terms)
    np.testing.assert_array_equal(a.counts, b.counts)
    assert a.terms.min() >= 0 and a.terms.max() < a.V and a.counts.min() >= 1
    # no duplicate term inside a document (the reference's scatter is last-write-wins on duplicates, LDA.jl:131)
    key = np.repeat(np.arange(a.M), np.diff(a.N_cumsum)) * a.V + a.terms
    assert len(np.unique(key)) == len(key)
    c = tm.synth.citeu_shaped(M=500, V=800, U=300)
    assert c.readers.max() < c.U and c.R_cumsum[-1] == len(c.readers)


def test_model_construction_mirrors_gpuLDA(tm):
    from scipy.special import digamma

    c = tm.synth.gencorp_lda(M=12, V=40, K=3, seed=0)
    m = tm.gpuLDA(tm.Corpus.from_csr(c), 4, seed=0)
    assert (m.K, m.M, m.V) == (4, 12, c.V)
    np.testing.assert_array_equal(m.N, np.diff(c.N_cumsum))
    assert m.C.sum() == c.counts.sum()
    np.testing.assert_allclose(m.alpha, 1.0)                                            # gpuLDA.jl:55
    np.testing.assert_allclose(m.beta.sum(axis=1), 1.0, rtol=1e-5)                      # gpuLDA.jl:56
    np.testing.assert_allclose(m.Elogtheta, -(np.euler_gamma + digamma(4)), rtol=1e-6)  # gpuLDA.jl:57
    np.testing.assert_allclose(m.gamma, 1.0)                                            # gpuLDA.jl:60
    assert m.beta.flags["F_CONTIGUOUS"] and m.beta.dtype == np.float32
    tm.check_model(m)
    with pytest.raises(ValueError):
        tm.gpuLDA(tm.Corpus.from_csr(c), 0)


def test_train_argument_validation_needs_no_device(tm):
    c = tm.synth.gencorp_lda(M=12, V=40, K=3, seed=0)
    m = tm.gpuLDA(tm.Corpus.from_csr(c), 4, seed=0)
    for kw in (dict(tol=-1.0), dict(ntol=-1.0), dict(vtol=-0.5), dict(iter=-1), dict(niter=-2), dict(viter=-1),
               dict(checkelbo=0), dict(checkelbo=2.5)):
        with pytest.raises(ValueError):
            tm.train(m, printelbo=False, **kw)                  # gpuLDA.jl:349-351
    m.alpha = np.array([1, 1, -1, 1], dtype=np.float32)
    with pytest.raises(tm.TopicModelError):
        tm.train(m, printelbo=False)                            # modelutils.jl:263
    m = tm.gpuLDA(tm.Corpus.from_csr(c), 4, seed=0)
    m.beta = m.beta * 2
    with pytest.raises(tm.TopicModelError):
        tm.train(m, printelbo=False)                            # modelutils.jl:265


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, K, out):
    import torch
    import torch.distributed as dist

    import oracle
    import topicmodelsvb_b200.synth as synth

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = synth.gencorp_lda(M=90, V=200, K=4, seed=3)
    sh = c.shard(rank, world)
    beta0 = synth.init_beta(K, c.V, seed=7)
    st = oracle.LDAState(K, sh.M, sh.V, beta=beta0)
    elbos = []
    for it in range(3):
        # E-step on the shard, then ONE sum over ranks of [stats | sum_d Elogtheta]
        stats, _ = oracle.lda_estep(st, sh.N_cumsum, sh.terms, sh.counts)
        t_stats = torch.from_numpy(stats)
        t_small = torch.from_numpy(st.Elogtheta.sum(axis=0))
        dist.all_reduce(t_stats)
        dist.all_reduce(t_small)
        st.beta_old = st.beta.copy()
        st.beta = stats / stats.sum(axis=0, keepdims=True)        # replicated M-step on the reduced buffer
        st.alpha, _ = oracle.lda_update_alpha(K, c.M, st.alpha, t_small.numpy())   # M_total, not the shard size
        e = torch.tensor([oracle.lda_elbo(st, sh.N_cumsum, sh.terms, sh.counts)], dtype=torch.float64)
        dist.all_reduce(e)
        elbos.append(float(e))
    if rank == 0:
        np.savez(out, elbo=np.array(elbos), beta=st.beta, alpha=st.alpha)
    dist.destroy_process_group()


def test_two_rank_gloo_exchange_matches_single_process(tm, orc, tmp_path):
    """The N > 1 protocol (shard d % N, all-reduce statistics, replicated M-step, alpha with M_total)
    reproduces the single-process trajectory; run on CPU with the oracle as the per-rank engine."""
    import torch.multiprocessing as mp

    K, world = 5, 2
    out = str(tmp_path / "r0.npz")
    mp.spawn(_rank_main, args=(world, _free_port(), K, out), nprocs=world, join=True)
    got = np.load(out)
    c = tm.synth.gencorp_lda(M=90, V=200, K=4, seed=3)
    st = orc.LDAState(K, c.M, c.V, beta=tm.synth.init_beta(K, c.V, seed=7))
    tr, _, _ = orc.lda_train(st, c.N_cumsum, c.terms, c.counts, iter=3, tol=0.0)
    np.testing.assert_allclose(got["elbo"], tr[1:], rtol=1e-11)
    np.testing.assert_allclose(got["beta"], st.beta, rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(got["alpha"], st.alpha, rtol=1e-9)


def _handshake_main(rank, world, port, fail_rank, out_dir):
    import json

    import torch.distributed as dist

    import topicmodelsvb_b200 as tm

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    red = tm.dist.Reducer()
    seen = {}

    def export(buf, n):                      # stands in for tmvb_lda_comm_export: a recognisable per-rank blob
        assert n == 512
        buf.raw = bytes([rank + 1]) * n

    def connect(r, w, blobs, n):             # stands in for tmvb_lda_comm_connect
        if rank == fail_rank:
            raise RuntimeError("this rank cannot map its peers")
        seen.update(rank=r, world=w, blobs=[blobs[i * n] for i in range(w)], size=len(blobs))

    ok = red.connect_peers(export, connect, 512)
    with open(os.path.join(out_dir, "r%d.json" % rank), "w") as f:
        json.dump({"ok": ok, "seen": seen}, f)
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 1])
def test_peer_handshake_protocol_two_ranks(tm, tmp_path, fail_rank):
    """The host side of the peer-memory exchange (dist.Reducer.connect_peers): every rank receives all blobs in rank
    order, and if ANY rank cannot map its peers every rank falls back (to the NCCL all-reduce path) together."""
    import json

    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_handshake_main, args=(world, _free_port(), fail_rank, str(tmp_path)), nprocs=world, join=True)
    res = [json.load(open(tmp_path / ("r%d.json" % r))) for r in range(world)]
    assert all(r["ok"] == (fail_rank < 0) for r in res)
    for r, x in enumerate(res):
        if r != fail_rank:
            assert x["seen"]["rank"] == r and x["seen"]["world"] == world and x["seen"]["size"] == world * 512
            assert x["seen"]["blobs"] == [1, 2]


def test_ctpf_recommendations_follow_the_reference_loops(tm):
    """libs / drecs / urecs (gpuCTPF.jl:88-91, 709-731) of the host mirror == the reference's loops over the dense score matrix
    (findall(mask)[reverse(sortperm(scores[mask]))], ties included); no device needed."""
    c = tm.synth.gencorp_ctpf(M=30, V=80, U=17, K=3, seed=5)
    K = 4
    m = tm.gpuCTPF(tm.Corpus.from_csr(c), K, seed=1)
    rng = np.random.default_rng(0)
    for n, cols in (("he", c.U), ("gimel", c.M), ("zayin", c.M)):
        setattr(m, n, np.asfortranarray(rng.integers(1, 4, size=(K, cols)).astype(np.float32)))   # small integers: many tied scores
    for n in ("vav", "dalet", "het"):
        setattr(m, n, rng.integers(1, 3, size=K).astype(np.float32))
    scores = m.scores()
    assert scores.shape == (c.M, c.U)
    libs = [[d + 1 for d in range(c.M) if u in c.readers[c.R_cumsum[d]:c.R_cumsum[d + 1]]] for u in range(c.U)]
    for u in range(c.U):
        np.testing.assert_array_equal(m.libs[u], libs[u])
        ur = np.ones(c.M, bool)
        ur[np.array(libs[u], dtype=int) - 1] = False
        idx = np.flatnonzero(ur)
        want = idx[np.argsort(scores[idx, u], kind="stable")[::-1]] + 1
        np.testing.assert_array_equal(m.urecs[u], want)
    for d in range(c.M):
        nr = np.ones(c.U, bool)
        nr[c.readers[c.R_cumsum[d]:c.R_cumsum[d + 1]]] = False
        idx = np.flatnonzero(nr)
        want = idx[np.argsort(scores[d, idx], kind="stable")[::-1]] + 1
        np.testing.assert_array_equal(m.drecs[d], want)
    assert len(m.drecs) == c.M and len(m.urecs) == c.U


def test_filtered_model_constructors_and_check_model(tm):
    """gpufLDA / gpufCTM host mirrors (fields of fLDA.jl:6-58 / fCTM.jl:6-64) and check_model(::fLDA) (modelutils.jl:69-98) without a
    device: initial state, shapes, the keyword validation of train!, and that train! on an all-empty corpus returns without
    touching the library (fLDA.jl:219)."""
    import pytest

    c = tm.synth.gencorp_lda(M=12, V=40, K=3, seed=2)
    m = tm.gpufLDA(tm.Corpus.from_csr(c), 4, seed=0)
    assert m.eta == 0.5 and m.alpha.shape == (4,) and m.kappa.shape == (c.V,) and abs(float(m.kappa.sum()) - 1) < 1e-5
    assert m.beta.shape == (4, c.V) and m.Elogtheta.shape == (4, c.M) and m.gamma.shape == (4, c.M)
    assert m.tau.shape == (len(c.terms),) and np.all(m.tau == 0.5) and np.all(m.tau_old == 0.5)
    np.testing.assert_array_equal(m.tau_of(3), m.tau[c.N_cumsum[3]:c.N_cumsum[4]])
    np.testing.assert_allclose(m.beta.sum(axis=1), 1.0, rtol=1e-5)
    tm.check_model(m)
    m.eta = 1.2
    with pytest.raises(tm.TopicModelError, match="eta must belong"):
        tm.check_model(m)
    m.eta = 0.5
    m.alpha = np.array([1, 1, -1, 1], np.float32)
    with pytest.raises(tm.TopicModelError, match="alpha must be positive"):
        tm.check_model(m)
    m.alpha = np.ones(4, np.float32)
    m.kappa = np.full(c.V, 1.0, np.float32)
    with pytest.raises(tm.TopicModelError, match="kappa must be a probability vector"):
        tm.check_model(m)
    m = tm.gpufLDA(tm.Corpus.from_csr(c), 4, seed=0)
    for kw in (dict(tol=-1.0), dict(iter=-1), dict(checkelbo=0), dict(viter=0)):
        with pytest.raises(ValueError):
            tm.train(m, printelbo=False, **kw)
    with pytest.raises(ValueError, match="positive integer"):
        tm.gpufLDA(tm.Corpus.from_csr(c), 0)
    fc = tm.gpufCTM(tm.Corpus.from_csr(c), 4, seed=0)
    assert isinstance(fc, tm.gpuCTM) and fc.eta == 0.5 and fc.kappa.shape == (c.V,) and fc.tau.shape == (len(c.terms),)
    assert fc.lam.shape == (4, c.M) and np.all(fc.vsq == 1) and np.all(fc.logzeta == 0.5)
    tm.check_model(fc)
    # an all-empty corpus: iter = 0, nothing is uploaded (no device is needed)
    empty = tm.synth.CSR(3, 10, np.zeros(4, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64))
    me = tm.gpufLDA(tm.Corpus.from_csr(empty), 2)
    tm.train(me, iter=5, printelbo=False)
    assert me.elbo == 0.0 and [len(t) for t in me.topics] == [10, 10]


def test_connect_model_peers_single_rank_is_a_no_op(tm):
    """dist.connect_model_peers: no reducer / one rank / TMVB_P2P=0 keep the NCCL path without touching the library."""
    class _M:
        reducer, _h = None, None
    assert tm.dist.connect_model_peers(_M(), "ctm") is False


def test_document_lengths_are_checked_by_identity_or_value(tm):
    """check_model's "N must contain document lengths." (modelutils.jl:258): model.N is the corpus' cached read-only length vector,
    so the check is an identity test; a rebound N is compared by value, and writing into the shared vector is refused."""
    from topicmodelsvb_b200 import gpu_lda
    c = tm.synth.gencorp_lda(M=30, V=60, K=3, seed=2)
    corp = tm.Corpus.from_csr(c)
    model = tm.gpuLDA(corp, 4)
    assert model.N is model.corp.lengths() and model.N.dtype == np.int64
    np.testing.assert_array_equal(model.N, np.diff(c.N_cumsum))
    gpu_lda.check_model(model)
    with pytest.raises(ValueError):
        model.N[0] += 1                              # the vector is shared with the corpus: read-only
    model.N = np.array(model.N)                      # an equal copy passes by value
    gpu_lda.check_model(model)
    model.N = model.N + 1
    with pytest.raises(tm._lib.TopicModelError, match="N must contain document lengths"):
        gpu_lda.check_model(model)
    # a corpus of Document objects: one vector per flattening
    docs = [tm.Document(terms=np.array([1, 2, 3]), counts=np.array([1, 1, 2])), tm.Document(terms=np.array([2]), counts=np.array([4]))]
    m2 = tm.gpuLDA(tm.Corpus(docs, vocab=5), 2)
    np.testing.assert_array_equal(m2.N, [3, 1])
    gpu_lda.check_model(m2)
