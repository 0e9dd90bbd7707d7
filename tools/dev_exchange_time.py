"""Duration of the exchange kernel (GPU box, torchrun): NSF LDA K=50 doc-sharded over WORLD ranks with TMVB_UNFUSED=1, so that
tmvb_lda_exchange_mstep is a call of its own bracketed by the handle's CUDA events (stats().mstep_ms); the E-step beside it.
The kernel cannot be captured by ncu (one profiled rank of a peer-synchronised kernel dead-locks the replay)."""
import os, sys
os.environ["TMVB_UNFUSED"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch, torch.distributed as dist
import bench
import topicmodelsvb_b200 as tm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tm.build()
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "nsf_lda_k50"]
ws = torch.cuda.Stream(); torch.cuda.set_stream(ws)
red = tm.dist.Reducer()
shard, M_total, nnz, V, U, desc, _ = bench.load_corpus(tm.synth, cfg, rank, world, "auto")
arm = bench.Arm(tm, cfg, shard, V, U, M_total, red, ws.cuda_stream)
m = arm.model
m.update_buffer()
K = cfg["K"]
es, ms = [], []
for it in range(12):
    dist.barrier(); torch.cuda.synchronize()
    m.estep(10, 1.0 / K**2, want_elbo=True)
    m.update_beta()                       # tmvb_lda_exchange_mstep
    m.update_alpha(1000, 1.0 / K**2)
    e = m.update_elbo(0)
    st = m.stats()
    if it >= 2:
        es.append(st.estep_ms); ms.append(st.mstep_ms)
t = torch.tensor([np.mean(es), np.mean(ms), np.min(ms), np.max(ms)], dtype=torch.float64, device="cuda")
out = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(out, t)
if rank == 0:
    print("world %d p2p %s: per rank [estep_ms mean, exchange_ms mean, min, max] over 10 iterations" % (world, getattr(m, "_p2p", None)))
    for r, o in enumerate(out):
        print("  rank %d  %.4f  %.4f  %.4f  %.4f" % (r, *o.tolist()))
dist.destroy_process_group()
