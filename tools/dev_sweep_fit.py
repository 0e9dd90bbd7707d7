"""Developer probe (GPU box): E-step time as a function of the number of sweeps (vtol = 0 forces exactly viter sweeps)
-> per-document fixed cost vs per-sweep cost of the LDA E-step kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import topicmodelsvb_b200 as tm

tm.build()
K = int(os.environ.get("K", 50))
c = tm.synth.load_packed("nsf") or tm.synth.nsf_shaped()
model = tm.gpuLDA(tm.Corpus.from_csr(c), K, seed=7)
model.update_buffer()
for viter in [int(x) for x in os.environ.get("VITERS", "1,2,4,8,10").split(",")]:
    ts = []
    for rep in range(3):
        model.update_buffer()          # same state every time
        model.estep(viter, 0.0, want_elbo=False)
        model.update_beta()
        ts.append(model.stats().estep_ms)
    print("viter %2d estep_ms %s" % (viter, " ".join("%.3f" % t for t in ts)), flush=True)
