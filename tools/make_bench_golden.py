"""Full-size oracle ELBO traces for the four bench configurations -> tests/golden/bench_traces.json.

bench.py prints `parity: {max_rel_vs_oracle, iters}` in every line by comparing the ELBO after every outer iteration
of one `train(iter=N)` call with these committed traces (bench.py itself never runs the oracle outside its
cpu_baseline / --impl reference legs), and the full-size GPU parity tests read the same file instead of spending
minutes of CPU time on the GPU box.

The traces are the fp64 C oracle's (`oracle/*_oracle.c`, a restatement of LDA.jl / CTM.jl / CTPF.jl train!): PARITY
UNPINNED with respect to the reference itself (no Julia here; see tools/reference_golden.jl for the recipe that pins it).

usage: python tools/make_bench_golden.py [nsf_lda_k50] [citeu_ctm_k30] [citeu_ctpf_k30] [synth_lda_k200]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import topicmodelsvb_b200.synth as synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "bench_traces.json")
NT = oracle.host_threads()


def nsf_lda_k50():
    c = synth.load_packed("nsf")
    K = 50
    beta0 = synth.init_beta(K, c.V, seed=7).astype(np.float32)
    st = oracle.LDAState(K, c.M, c.V, beta=beta0)
    t0 = time.time()
    trace, sweeps, _ = oracle.lda_train(st, c.N_cumsum, c.terms, c.counts, iter=10, tol=0.0, viter=10, checkelbo=1, nthreads=NT)
    return dict(corpus="data/_packed/nsf.npz", M=c.M, V=c.V, K=K, nnz=c.nnz, init="synth.init_beta(K, V, seed=7) as float32", iter=10,
                viter=10, elbo=trace.tolist(), sweeps=sweeps.tolist(), alpha=st.alpha.tolist(), seconds=time.time() - t0, threads=NT)


def citeu_ctm_k30():
    c = synth.load_packed("citeu")
    K = 30
    beta0 = synth.init_beta(K, c.V, seed=7).astype(np.float32)
    st = oracle.CTMState(K, c.M, c.V, beta0)
    t0 = time.time()
    trace, sweeps, _ = oracle.ctm_train(st, c.N_cumsum, c.terms, c.counts, iter=10, tol=0.0, viter=10, checkelbo=1, nthreads=NT)
    return dict(corpus="data/_packed/citeu.npz", M=c.M, V=c.V, K=K, nnz=c.nnz, init="synth.init_beta(K, V, seed=7) as float32", iter=10,
                viter=10, elbo=trace.tolist(), sweeps=sweeps.tolist(), mu=st.mu.tolist(), seconds=time.time() - t0, threads=NT)


def citeu_ctpf_k30():
    c = synth.load_packed("citeu")
    K = 30
    alef0 = synth.init_alef(K, c.V, seed=7).astype(np.float32)
    st = oracle.CTPFState(K, c.M, c.V, c.U, alef0)
    t0 = time.time()
    trace, sweeps, _ = oracle.ctpf_train(st, c, iter=10, tol=0.0, viter=10, checkelbo=1, nthreads=NT)
    return dict(corpus="data/_packed/citeu.npz", M=c.M, V=c.V, U=c.U, K=K, nnz=c.nnz, init="synth.init_alef(K, V, seed=7) as float32",
                iter=10, viter=10, elbo=trace.tolist(), sweeps=sweeps.tolist(), bet=st.bet.tolist(), seconds=time.time() - t0, threads=NT)


def synth_lda_k200():
    """The first two 12 500-document blocks of cfg4 (the full million costs the fp64 oracle ~10 minutes per iteration)."""
    K, blocks = 200, 2
    c = synth.cfg4_shard(0, 1, M=blocks * synth.CFG4_BLOCK)
    beta0 = synth.init_beta(K, c.V, seed=7).astype(np.float32)
    st = oracle.LDAState(K, c.M, c.V, beta=beta0)
    t0 = time.time()
    trace, sweeps, _ = oracle.lda_train(st, c.N_cumsum, c.terms, c.counts, iter=5, tol=0.0, viter=10, checkelbo=1, nthreads=NT)
    return dict(corpus="synth.cfg4_shard(0, 1, M=25000)  (first two blocks of cfg4)", M=c.M, V=c.V, K=K, nnz=c.nnz,
                init="synth.init_beta(K, V, seed=7) as float32", iter=5, viter=10, elbo=trace.tolist(), sweeps=sweeps.tolist(),
                alpha_head=st.alpha[:8].tolist(), seconds=time.time() - t0, threads=NT)


def _kappa0(V):
    return np.random.default_rng(8).dirichlet(np.ones(V)).astype(np.float32)      # bench.filtered_kappa0


def nsf_flda_k50():
    c = synth.load_packed("nsf")
    K = 50
    beta0 = synth.init_beta(K, c.V, seed=7).astype(np.float32)
    st = oracle.FLDAState(K, c.M, c.V, len(c.terms), beta0, _kappa0(c.V))
    t0 = time.time()
    trace, sweeps, _ = oracle.flda_train(st, c.N_cumsum, c.terms, c.counts, iter=5, tol=0.0, viter=10, checkelbo=1, nthreads=NT)
    return dict(corpus="data/_packed/nsf.npz", M=c.M, V=c.V, K=K, nnz=c.nnz, init="synth.init_beta(K, V, seed=7), default_rng(8).dirichlet(ones(V)) as float32",
                iter=5, viter=10, elbo=trace.tolist(), sweeps=sweeps.tolist(), eta=float(st.eta[0]), seconds=time.time() - t0, threads=NT)


def citeu_fctm_k30():
    c = synth.load_packed("citeu")
    K = 30
    beta0 = synth.init_beta(K, c.V, seed=7).astype(np.float32)
    st = oracle.FCTMState(K, c.M, c.V, len(c.terms), beta0, _kappa0(c.V))
    t0 = time.time()
    trace, sweeps, _ = oracle.fctm_train(st, c.N_cumsum, c.terms, c.counts, iter=5, tol=0.0, viter=10, checkelbo=1, nthreads=NT)
    return dict(corpus="data/_packed/citeu.npz", M=c.M, V=c.V, K=K, nnz=c.nnz, init="synth.init_beta(K, V, seed=7), default_rng(8).dirichlet(ones(V)) as float32",
                iter=5, viter=10, elbo=trace.tolist(), sweeps=sweeps.tolist(), mu=st.mu.tolist(), seconds=time.time() - t0, threads=NT)


if __name__ == "__main__":
    which = sys.argv[1:] or ["nsf_lda_k50", "citeu_ctm_k30", "citeu_ctpf_k30", "synth_lda_k200"]
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for w in which:
        out[w] = globals()[w]()
        print(w, "%.1f s" % out[w]["seconds"], out[w]["elbo"][:3], "...", out[w]["elbo"][-1], flush=True)
        json.dump(out, open(OUT, "w"), indent=1)
