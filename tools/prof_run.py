"""Profiling driver (GPU box, under ncu --profile-from-start off): one bench configuration, `--warm` untimed outer iterations, then
`--iters` iterations between cudaProfilerStart/Stop.  TMVB_GRAPH=0 makes every bucket launch a plain kernel launch.

    ncu --profile-from-start off --metrics gpu__time_duration.sum ... python tools/prof_run.py --config nsf_lda_k50
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import topicmodelsvb_b200 as tm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="nsf_lda_k50")
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--e2e", action="store_true", help="profile one train(iter=1) call instead of resident iterations")
args = ap.parse_args()

tm.build()
cfg = bench.CONFIGS[args.config]
torch.cuda.set_device(0)
ws = torch.cuda.Stream()
torch.cuda.set_stream(ws)
shard, M_total, nnz, V, U, desc, _ = bench.load_corpus(tm.synth, cfg, 0, 1, "auto")
arm = bench.Arm(tm, cfg, shard, V, U, M_total, None, ws.cuda_stream)
rt = torch.cuda.cudart()
if args.e2e:
    arm.train(1)
    arm.reinit_host()
    torch.cuda.synchronize()
    rt.cudaProfilerStart()
    arm.train(1)
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
else:
    arm.model.update_buffer()
    for _ in range(args.warm):
        arm.step()
    torch.cuda.synchronize()
    rt.cudaProfilerStart()
    for _ in range(args.iters):
        e = arm.step()
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
    st = arm.model.stats()
    print(args.config, "M", M_total, "estep_ms %.3f mstep_ms %.3f sweeps/doc %.2f elbo %.8e" % (st.estep_ms, st.mstep_ms, st.sweeps / M_total, e))
