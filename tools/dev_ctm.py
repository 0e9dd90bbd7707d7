"""Developer probe (GPU box): gpuCTM / gpuCTPF E-step time on CiteULike for a few tile-capacity limits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import topicmodelsvb_b200 as tm
tm.build()
c = tm.synth.load_packed("citeu") or tm.synth.citeu_shaped()
K = 30
for which in sys.argv[1:] or ["ctm"]:
    for cap in os.environ.get("CAPS", "0,16,32,64").split(","):
        if cap == "0":
            os.environ.pop("TMVB_TILE_CAP_MAX", None)
        else:
            os.environ["TMVB_TILE_CAP_MAX"] = cap
        if which == "ctm":
            m = tm.gpuCTM(tm.Corpus.from_csr(tm.synth.CSR(c.M, c.V, c.N_cumsum, c.terms, c.counts)), K, seed=7)
            m.beta = np.asfortranarray(tm.synth.init_beta(K, c.V, seed=7).T.astype(np.float32))
        else:
            m = tm.gpuCTPF(tm.Corpus.from_csr(c), K, seed=7)
            m.alef = np.asfortranarray(tm.synth.init_alef(K, c.V, seed=7).T.astype(np.float32))
        m.update_buffer()
        out = []
        for it in range(4):
            if which == "ctm":
                m.estep(1000, 1.0 / K**2, 10, 1.0 / K**2, want_elbo=True)
            else:
                m.estep(10, 1.0 / K**2, want_elbo=True)
            m.mstep()
            e = m.update_elbo(0)
            out.append((m.stats().estep_ms, e))
        print(which, "cap_max", cap, " ".join("%.3f" % o[0] for o in out), "| elbo", " ".join("%.7e" % o[1] for o in out), flush=True)
        m.close()
