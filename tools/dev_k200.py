"""Developer probe (GPU box): BASELINE configs[4]-shaped run, gpuLDA K=200 on a synthetic M x 50k-vocab corpus (tile kernel)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import topicmodelsvb_b200 as tm

tm.build()
K, M, V = 200, int(os.environ.get("M", 200000)), 50000
t = time.time()
c = tm.synth.nsf_shaped(M=M, V=V, seed=1)
print("corpus", c.M, c.V, c.nnz, "gen s %.1f" % (time.time() - t), flush=True)
model = tm.gpuLDA(tm.Corpus.from_csr(c), K, seed=7)
model.update_buffer()
for it in range(4):
    model.estep(10, 1.0 / K**2, want_elbo=True)
    model.update_beta()
    model.update_alpha(1000, 1.0 / K**2)
    st = model.stats()
    alg = c.nnz * (8 * K + 8) + 12 * K * c.M
    print(it, "estep_ms %.3f mstep_ms %.3f sweeps/doc %.2f docs/s %.3e  alg GB/s %.0f  elbo %.6e" % (
        st.estep_ms, st.mstep_ms, st.sweeps / c.M, c.M / (st.estep_ms * 1e-3), alg / (st.estep_ms * 1e-3) / 1e9, model.update_elbo(0)), flush=True)
