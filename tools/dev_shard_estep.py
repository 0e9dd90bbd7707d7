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
out = []
for it in range(int(os.environ.get("ITERS", 6))):
    model.iterate(10, 1.0 / K**2, 1000, 1.0 / K**2, want_elbo=True)
    st = model.stats()
    out.append(st.estep_ms)
print("WORLD", world, "streams", os.environ.get("TMVB_STREAMS", "4"), "classes", os.environ.get("TMVB_HYB_CLASSES", "default"), "docs", c.M,
      "estep_ms", " ".join("%.3f" % x for x in out), flush=True)
