import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import topicmodelsvb_b200 as tm
tm.build()
K = 50
c = tm.synth.load_packed("nsf") or tm.synth.nsf_shaped()
pin = tm._lib.pinned_copy
c = c._replace(N_cumsum=pin(c.N_cumsum), terms=pin(c.terms), counts=pin(c.counts))
model = tm.gpuLDA(tm.Corpus.from_csr(c), K, seed=7)
tm.train(model, iter=1, tol=0.0, printelbo=False)
tm.train(model, iter=1, tol=0.0, printelbo=False)
pr = cProfile.Profile(); pr.enable()
t = time.perf_counter()
for _ in range(3):
    tm.train(model, iter=1, tol=0.0, printelbo=False)
print("per call ms", (time.perf_counter() - t) / 3 * 1e3)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
