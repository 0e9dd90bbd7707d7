"""cProfile of one train(model, iter=1) call of a bench configuration (GPU box): usage python tools/dev_e2e_profile2.py <config>"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import topicmodelsvb_b200 as tm
tm.build()
name = sys.argv[1]
cfg = bench.CONFIGS[name]
torch.cuda.set_device(0)
ws = torch.cuda.Stream(); torch.cuda.set_stream(ws)
shard, M_total, nnz, V, U, desc, _ = bench.load_corpus(tm.synth, cfg, 0, 1, "auto")
pin = tm._lib.pinned_copy
shard = shard._replace(N_cumsum=pin(shard.N_cumsum), terms=pin(shard.terms), counts=pin(shard.counts))
arm = bench.Arm(tm, cfg, shard, V, U, M_total, None, ws.cuda_stream)
for _ in range(2):
    arm.reinit_host(); arm.train(1)
pr = cProfile.Profile(); pr.enable()
t = time.perf_counter()
for _ in range(3):
    arm.train(1)
print(name, "per call ms", (time.perf_counter() - t) / 3 * 1e3)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(16)
