"""GPU tests of the 'next' rows of SURVEY 8(f): predict (E-step-only inference) and the 2-GPU NCCL exchange."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_predict_lda_matches_oracle_estep(tm, orc):
    """predict(corp, train_model) (modelutils.jl:831-855): frozen alpha/beta, per-document sweeps on unseen documents."""
    train = tm.synth.gencorp_lda(M=150, V=400, K=5, seed=1)
    new = tm.synth.gencorp_lda(M=40, V=400, K=5, seed=2)
    K = 6
    m = tm.gpuLDA(tm.Corpus.from_csr(train), K, seed=3)
    tm.train(m, iter=5, tol=0.0, printelbo=False)
    p = tm.predict(tm.Corpus.from_csr(new), m)
    st = orc.LDAState(K, new.M, new.V, alpha=m.alpha, beta=np.ascontiguousarray(m.beta.T))
    orc.lda_estep(st, new.N_cumsum, new.terms, new.counts, viter=10, want_stats=False)
    np.testing.assert_allclose(p.gamma.T, st.gamma, rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(p.Elogtheta.T, st.Elogtheta, rtol=2e-3, atol=5e-4)
    td = tm.topicdist(p, 0)
    np.testing.assert_allclose(td, st.gamma[0] / st.gamma[0].sum(), rtol=2e-3, atol=1e-6)
    with pytest.raises(tm.CorpusError):
        tm.predict(tm.Corpus.from_csr(tm.synth.gencorp_lda(M=5, V=77, K=3, seed=0)), m)


def test_predict_ctm_matches_oracle_estep(tm):
    """predict(corp, train_model::gpuCTM) (modelutils.jl:886-913) == the per-document loop of the CPU model (update_phi!,
    update_logzeta!, update_vsq!, update_lambda!; CTM.jl:129-178) run by the NumPy twin of the oracle with the trained mu, sigma,
    beta frozen."""
    from oracle.numpy_twin import CTMTwin

    train = tm.synth.gencorp_lda(M=100, V=300, K=4, seed=1)
    new = tm.synth.gencorp_lda(M=30, V=300, K=4, seed=5)
    # CTM has no epsilon in phi (CTM.jl:177): a term unseen in training has beta = 0 for every topic and yields NaN in the
    # reference as well, so keep only terms the training corpus contains
    seen = np.zeros(300, bool)
    seen[train.terms] = True
    keep = seen[new.terms]
    doc = np.repeat(np.arange(new.M), np.diff(new.N_cumsum))[keep]
    off = np.concatenate([[0], np.cumsum(np.bincount(doc, minlength=new.M))]).astype(np.int64)
    new = tm.synth.CSR(new.M, new.V, off, new.terms[keep], new.counts[keep])
    K = 5
    m = tm.gpuCTM(tm.Corpus.from_csr(train), K, seed=3)
    tm.train(m, iter=3, tol=0.0, printelbo=False)
    p = tm.predict(tm.Corpus.from_csr(new), m)
    assert p.lam.shape == (K, new.M) and np.all(np.isfinite(p.lam)) and np.all(p.vsq > 0)
    td = tm.topicdist(p, 3)
    assert abs(td.sum() - 1) < 1e-6
    # the oracle's loop on the new documents with the device's globals
    tw = CTMTwin(new.N_cumsum, new.terms, new.counts, K, new.V, np.asarray(m.beta, dtype=np.float64).T)
    tw.mu = np.asarray(m.mu, dtype=np.float64)
    tw.sigma = np.asarray(m.sigma, dtype=np.float64)
    tw.invsigma = np.linalg.inv(tw.sigma)
    tol = ntol = 1.0 / K**2
    for d in range(new.M):
        if new.N_cumsum[d + 1] == new.N_cumsum[d]:
            continue
        for _ in range(10):
            tw.update_phi(d)
            tw.update_logzeta(d)
            tw.update_vsq(d, 1000, ntol)
            tw.update_lambda(d, 1000, ntol)
            if np.linalg.norm(tw.lam[d] - tw.lam_old[d]) < tol:
                break
    nz = np.diff(new.N_cumsum) > 0
    np.testing.assert_allclose(p.lam.T[nz], tw.lam[nz], rtol=5e-3, atol=5e-3)
    np.testing.assert_allclose(p.vsq.T[nz], tw.vsq[nz], rtol=5e-3, atol=1e-5)
    np.testing.assert_allclose(p.logzeta[nz], tw.logzeta[nz], rtol=2e-3, atol=2e-3)
    # predict scatters nothing: the statistics buffer of the predict handle is still zero
    import ctypes as C
    sp, sn = C.c_void_p(), C.c_int64()
    tm._lib.check(tm._lib.load().tmvb_ctm_reduce_buffers(p._handle(), C.byref(sp), C.byref(sn), None, None))
    import torch
    stats = torch.as_tensor(tm.dist._DevBuf(sp.value, sn.value, "<f4"), device="cuda")
    assert float(stats.abs().sum()) == 0.0


_WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import topicmodelsvb_b200 as tm
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
work = torch.cuda.Stream(); torch.cuda.set_stream(work)
red = tm.dist.Reducer()
c = tm.synth.gencorp_lda(M=400, V=600, K=6, seed=11)
K = 8
beta0 = tm.synth.init_beta(K, c.V, seed=7).astype(np.float32)
sh = c.shard(rank, world)
m = tm.gpuLDA(tm.Corpus.from_csr(sh), K, reducer=red, M_total=c.M, stream=work.cuda_stream)
m.beta = np.array(beta0.T, order="F", copy=True)
tr = []
tm.train(m, iter=4, tol=0.0, printelbo=False, trace=tr)
import ctypes
st = ctypes.c_int(0)
tm._lib.check(tm._lib.load().tmvb_lda_comm_status(m._handle(), ctypes.byref(st)))
bsum = torch.tensor([float(np.abs(m.beta).sum()), float(m.beta[1, 5])], dtype=torch.float64, device="cuda")
both = [torch.empty_like(bsum) for _ in range(world)]
dist.all_gather(both, bsum)
if rank == 0:
    print("RESULT " + json.dumps({"elbo": tr, "alpha": m.alpha.tolist(), "p2p": bool(m._p2p), "status": st.value,
                                  "beta_probe": [b.tolist() for b in both]}))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("p2p", ["1", "unfused", "0"])
def test_two_gpu_exchange_matches_single_gpu(tm, orc, tmp_path, p2p):
    """Doc-sharded d %% 2 over two GPUs == the single-process trajectory, with the statistics exchanged by the fused
    peer-memory kernel (tmvb_lda_exchange_mstep, TMVB_P2P=1) and by NCCL all-reduce + tmvb_lda_mstep (TMVB_P2P=0)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    # "1": fused iteration (exchange kernel with the update_alpha! CTA); "unfused": the exchange kernel between separate calls;
    # "0": NCCL all-reduce + tmvb_lda_mstep
    env = dict(os.environ, TMVB_P2P="0" if p2p == "0" else "1", TMVB_UNFUSED="1" if p2p == "unfused" else "0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", {"1": "29533", "unfused": "29535", "0": "29534"}[p2p], str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    got = json.loads(line[7:])
    assert got["status"] == 0
    assert got["p2p"] == (p2p != "0"), "the peer-memory exchange was not used"
    assert got["beta_probe"][0] == got["beta_probe"][1], "ranks disagree on beta"
    c = tm.synth.gencorp_lda(M=400, V=600, K=6, seed=11)
    K = 8
    beta0 = tm.synth.init_beta(K, c.V, seed=7).astype(np.float32)
    st = orc.LDAState(K, c.M, c.V, beta=beta0)
    ref, _, _ = orc.lda_train(st, c.N_cumsum, c.terms, c.counts, iter=4, tol=0.0)
    np.testing.assert_allclose(got["elbo"], ref, rtol=2e-6)
    np.testing.assert_allclose(got["alpha"], st.alpha, rtol=5e-4)


_WORKER_CTX = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import topicmodelsvb_b200 as tm
which = %r
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
work = torch.cuda.Stream(); torch.cuda.set_stream(work)
red = tm.dist.Reducer()
K = 7
if which == "ctm":
    c = tm.synth.gencorp_lda(M=300, V=500, K=5, seed=13)
    m = tm.gpuCTM(tm.Corpus.from_csr(c.shard(rank, world)), K, reducer=red, M_total=c.M, stream=work.cuda_stream)
    m.beta = np.array(tm.synth.init_beta(K, c.V, seed=7).astype(np.float32).T, order="F", copy=True)
else:
    c = tm.synth.gencorp_ctpf(M=300, V=500, U=90, K=5, seed=13)
    m = tm.gpuCTPF(tm.Corpus.from_csr(c.shard(rank, world)), K, reducer=red, M_total=c.M, stream=work.cuda_stream)
    m.alef = np.array(tm.synth.init_alef(K, c.V, seed=7).astype(np.float32).T, order="F", copy=True)
tr = []
tm.train(m, iter=4, tol=0.0, printelbo=False, trace=tr)
glob = m.beta if which == "ctm" else m.alef
probe = torch.tensor([float(np.abs(glob).sum()), float(glob[1, 5])], dtype=torch.float64, device="cuda")
both = [torch.empty_like(probe) for _ in range(world)]
dist.all_gather(both, probe)
if rank == 0:
    out = {"elbo": tr, "probe": [b.tolist() for b in both], "p2p": bool(getattr(m, "_p2p", False))}
    if which == "ctm":
        out.update(mu=np.asarray(m.mu, dtype=float).tolist(), sigma=np.asarray(m.sigma, dtype=float).tolist())
    else:
        out.update(vav=np.asarray(m.vav, dtype=float).tolist(), bet=np.asarray(m.bet, dtype=float).tolist(), he_sum=float(np.asarray(m.he, dtype=np.float64).sum()))
    print("RESULT " + json.dumps(out))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("p2p", ["1", "0"])
@pytest.mark.parametrize("which", ["ctm", "ctpf"])
def test_two_gpu_ctm_ctpf_match_oracle(tm, orc, tmp_path, which, p2p):
    """gpuCTM / gpuCTPF doc-sharded d %% 2 over two GPUs == the oracle's single-process trajectory (CTM.jl:185-217 /
    CTPF.jl:344-371), with the sufficient statistics summed by the one-kernel peer-memory all-reduce (tmvb_*_peer_reduce,
    TMVB_P2P=1) and by NCCL all-reduces of tmvb_*_reduce_buffers (TMVB_P2P=0)."""
    import json

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    script = tmp_path / "worker_ctx.py"
    script.write_text(_WORKER_CTX % (ROOT, which))
    port = {"ctm": 29541, "ctpf": 29543}[which] + int(p2p)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600, env=dict(os.environ, TMVB_P2P=p2p))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert got["p2p"] == (p2p == "1"), "the peer-memory all-reduce was not used"
    assert got["probe"][0] == got["probe"][1], "ranks disagree on the global parameters"
    K = 7
    if which == "ctm":
        c = tm.synth.gencorp_lda(M=300, V=500, K=5, seed=13)
        st = orc.CTMState(K, c.M, c.V, tm.synth.init_beta(K, c.V, seed=7).astype(np.float32))
        ref = orc.ctm_train(st, c.N_cumsum, c.terms, c.counts, iter=4, tol=0.0)[0]
        ref = np.asarray(ref)[np.isfinite(ref)]
        np.testing.assert_allclose(got["elbo"], ref, rtol=2e-5)
        np.testing.assert_allclose(got["mu"], st.mu, rtol=5e-3, atol=5e-4)
        np.testing.assert_allclose(got["sigma"], st.sigma, rtol=5e-3, atol=5e-4)
    else:
        c = tm.synth.gencorp_ctpf(M=300, V=500, U=90, K=5, seed=13)
        st = orc.CTPFState(K, c.M, c.V, c.U, tm.synth.init_alef(K, c.V, seed=7).astype(np.float32))
        ref = orc.ctpf_train(st, c, iter=4, tol=0.0)[0]
        ref = ref[np.isfinite(ref)]
        np.testing.assert_allclose(got["elbo"], ref, rtol=2e-5)
        np.testing.assert_allclose(got["vav"], st.vav, rtol=1e-3)
        np.testing.assert_allclose(got["bet"], st.bet, rtol=1e-3)
        np.testing.assert_allclose(got["he_sum"], st.he.sum(), rtol=1e-4)


_WORKER_FILT = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import topicmodelsvb_b200 as tm
which = %r
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
work = torch.cuda.Stream(); torch.cuda.set_stream(work)
red = tm.dist.Reducer()
K = 6
c = tm.synth.gencorp_lda(M=300, V=503, K=5, seed=17)          # V not a multiple of four: the padded kappa statistics
sh = c.shard(rank, world)
beta0 = tm.synth.init_beta(K, c.V, seed=7).astype(np.float32)
kappa0 = np.random.default_rng(8).dirichlet(np.ones(c.V)).astype(np.float32)
if which == "flda":
    m = tm.gpufLDA(tm.Corpus.from_csr(sh), K, reducer=red, M_total=c.M, C_total=float(np.asarray(c.counts).sum()), stream=work.cuda_stream)
else:
    m = tm.gpufCTM(tm.Corpus.from_csr(sh), K, reducer=red, M_total=c.M, stream=work.cuda_stream)
m.beta, m.kappa = np.array(beta0.T, order="F", copy=True), kappa0.copy()
tr = []
tm.train(m, iter=3, tol=0.0, printelbo=False, trace=tr)
probe = torch.tensor([float(np.abs(m.beta).sum()), float(m.kappa[7])], dtype=torch.float64, device="cuda")
both = [torch.empty_like(probe) for _ in range(world)]
dist.all_gather(both, probe)
if rank == 0:
    print("RESULT " + json.dumps({"elbo": tr, "probe": [b.tolist() for b in both], "p2p": bool(getattr(m, "_p2p", False)), "eta": float(m.eta),
                                  "kappa_head": np.asarray(m.kappa[:8], dtype=float).tolist()}))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("p2p", ["1", "0"])
@pytest.mark.parametrize("which", ["flda", "fctm"])
def test_two_gpu_filtered_models_match_oracle(tm, orc, tmp_path, which, p2p):
    """gpufLDA / gpufCTM doc-sharded over two GPUs == the oracle's single-process trajectory (fLDA.jl:214-247 / fCTM.jl:249-290):
    K x V statistics, the V-vector of update_kappa!, and the small fp64 sums (incl. update_eta!'s numerator) cross the ranks."""
    import json

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    script = tmp_path / "worker_filt.py"
    script.write_text(_WORKER_FILT % (ROOT, which))
    port = {"flda": 29551, "fctm": 29553}[which] + int(p2p)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600, env=dict(os.environ, TMVB_P2P=p2p))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert got["p2p"] == (p2p == "1")
    assert got["probe"][0] == got["probe"][1], "ranks disagree on the global parameters"
    K = 6
    c = tm.synth.gencorp_lda(M=300, V=503, K=5, seed=17)
    beta0 = tm.synth.init_beta(K, c.V, seed=7).astype(np.float32)
    kappa0 = np.random.default_rng(8).dirichlet(np.ones(c.V)).astype(np.float32)
    if which == "flda":
        st = orc.FLDAState(K, c.M, c.V, len(c.terms), beta0, kappa0)
        ref = orc.flda_train(st, c.N_cumsum, c.terms, c.counts, iter=3, tol=0.0)[0]
        assert abs(got["eta"] - st.eta[0]) < 1e-4
    else:
        st = orc.FCTMState(K, c.M, c.V, len(c.terms), beta0, kappa0)
        ref = orc.fctm_train(st, c.N_cumsum, c.terms, c.counts, iter=3, tol=0.0)[0]
    np.testing.assert_allclose(got["elbo"], ref[np.isfinite(ref)], rtol=2e-5)
    np.testing.assert_allclose(got["kappa_head"], st.kappa[:8], rtol=5e-3, atol=1e-7)
