"""Profiling driver for tmvb_ctpf_recs (GPU box, under ncu): CiteULike-size gpuCTPF K=30, one training iteration, then the
recommendation step between cudaProfilerStart/Stop."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import topicmodelsvb_b200 as tm  # noqa: E402

tm.build()
c = tm.synth.load_packed("citeu") or tm.synth.citeu_shaped()
model = tm.gpuCTPF(tm.Corpus.from_csr(c), 30, seed=3)
tm.train(model, iter=1, tol=0.0, printelbo=False)
model.update_recs()
torch.cuda.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
t = time.perf_counter()
model.update_recs()
torch.cuda.synchronize()
print("update_recs M=%d U=%d: %.1f ms wall (scores + urecs + drecs to the host)" % (c.M, c.U, (time.perf_counter() - t) * 1e3))
rt.cudaProfilerStop()
t = time.perf_counter()
sc = model.scores()
print("host NumPy scores only: %.1f ms" % ((time.perf_counter() - t) * 1e3))
