"""Developer probe (GPU box): E-step time of the hybrid kernel for several class lists (TMVB_HYB_CLASSES) against the
register-resident kernel, same corpus and initial state; the ELBO after every iteration must agree between variants.
usage: python tools/dev_hyb.py [nsf|k200] spec1 spec2 ...      ("reg" = TMVB_LDA_HYB=0)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import topicmodelsvb_b200 as tm

tm.build()
which = sys.argv[1]
specs = sys.argv[2:]
if which == "nsf":
    K, c = 50, tm.synth.load_packed("nsf") or tm.synth.nsf_shaped()
else:
    K, c = 200, tm.synth.cfg4_shard(0, 1, M=int(os.environ.get("M", 100000)))
beta0 = np.asfortranarray(tm.synth.init_beta(K, c.V, seed=7).T.astype(np.float32))
iters = int(os.environ.get("ITERS", 4))
for spec in specs:
    os.environ.pop("TMVB_HYB_CLASSES", None)
    os.environ["TMVB_LDA_HYB"] = "1"
    if spec == "reg":
        os.environ["TMVB_LDA_HYB"] = "0"
    elif spec != "default":
        os.environ["TMVB_HYB_CLASSES"] = spec
    model = tm.gpuLDA(tm.Corpus.from_csr(c), K)
    model.beta = beta0.copy(order="F")
    try:
        model.update_buffer()
        out = []
        for it in range(iters):
            model.estep(10, 1.0 / K**2, want_elbo=True)
            model.update_beta()
            model.update_alpha(1000, 1.0 / K**2)
            e = model.update_elbo(0)
            st = model.stats()
            out.append((st.estep_ms, st.sweeps / c.M, e))
        print("%-40s" % spec, " ".join("%.3f" % o[0] for o in out), "| sweeps %.2f" % out[-1][1], "| elbo", " ".join("%.7e" % o[2] for o in out), flush=True)
    except Exception as ex:
        print("%-40s FAILED %r" % (spec, ex), flush=True)
    model.close()
