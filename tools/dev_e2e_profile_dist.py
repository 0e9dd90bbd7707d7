"""cProfile of train(model, iter=1) on rank 0 of a torchrun job (NSF LDA K=50, doc-sharded): where the per-call host time goes."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
import topicmodelsvb_b200 as tm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tm.build()
cfg = bench.CONFIGS["nsf_lda_k50"]
ws = torch.cuda.Stream(); torch.cuda.set_stream(ws)
red = tm.dist.Reducer()
shard, M_total, nnz, V, U, desc, _ = bench.load_corpus(tm.synth, cfg, rank, world, "auto")
pin = tm._lib.pinned_copy
shard = shard._replace(N_cumsum=pin(shard.N_cumsum), terms=pin(shard.terms), counts=pin(shard.counts))
arm = bench.Arm(tm, cfg, shard, V, U, M_total, red, ws.cuda_stream)
for _ in range(3):
    arm.reinit_host(); arm.train(1)
dist.barrier(); torch.cuda.synchronize()
pr = cProfile.Profile()
if rank == 0: pr.enable()
t = time.perf_counter()
for _ in range(5):
    arm.train(1)
dt = (time.perf_counter() - t) / 5 * 1e3
if rank == 0:
    pr.disable()
    print("world", world, "per call ms", dt)
    pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
dist.destroy_process_group()
