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
from oracle.numpy_twin import LDATwin  # noqa: E402
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


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(os.path.join(out, "lda_cfg0.npz"), **lda_cfg0())
    for f in sorted(os.listdir(out)):
        print(f, os.path.getsize(os.path.join(out, f)), "bytes")
