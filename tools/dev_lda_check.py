"""Developer probe (GPU box): parity + timing of the LDA path on NSF-size input."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import topicmodelsvb_b200 as tm

tm.build()
K = int(os.environ.get("K", 50))
c = tm.synth.load_packed("nsf") or tm.synth.nsf_shaped()
print("corpus", c.M, c.V, c.nnz, flush=True)
model = tm.gpuLDA(tm.Corpus.from_csr(c), K, seed=7)
t = time.time(); model.update_buffer(); print("update_buffer s", time.time() - t, flush=True)
for it in range(int(os.environ.get("ITERS", 6))):
    for want in (False, True):
        if want and it % 2: continue
        model.estep(10, 1.0 / K**2, want_elbo=want)
        model.update_beta()
        model.update_alpha(1000, 1.0 / K**2)
        st = model.stats()
        print(it, "elbo" if want else "noelbo", "estep_ms %.3f mstep_ms %.3f sweeps/doc %.2f" % (st.estep_ms, st.mstep_ms, st.sweeps / c.M),
              ("elbo %.6e" % model.update_elbo(0)) if want else "", flush=True)
print("mode1 elbo %.6e" % model.update_elbo(1))
