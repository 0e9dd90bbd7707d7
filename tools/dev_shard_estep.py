"""Developer probe (GPU box, one GPU): E-step time of ONE rank's shard (documents d % WORLD == 0) of NSF."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import topicmodelsvb_b200 as tm

tm.build()
K = 50
world = int(os.environ.get("WORLD", 8))
c = (tm.synth.load_packed("nsf") or tm.synth.nsf_shaped()).shard(0, world)
model = tm.gpuLDA(tm.Corpus.from_csr(c), K, seed=7)
model.update_buffer()
for it in range(int(os.environ.get("ITERS", 4))):
    model.estep(10, 1.0 / K**2, want_elbo=True)
    model.update_beta()
    model.update_alpha(1000, 1.0 / K**2)
    st = model.stats()
    print(it, "docs", c.M, "estep_ms %.3f mstep_ms %.3f sweeps/doc %.2f" % (st.estep_ms, st.mstep_ms, st.sweeps / c.M), flush=True)
