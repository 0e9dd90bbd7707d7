#!/usr/bin/env python
"""bench.py -- LDA K=50 on NSF (BASELINE.json configs[1]): full VI iterations on 1..N B200s.

A "step" is ONE outer VI iteration over the whole corpus: fused E-step (all inner sweeps of every
document + scatter of the K x V statistics) -> M-step normalise (N > 1: one fused peer-memory kernel that
reduce-scatters the statistics over NVLink, normalises and all-gathers; NCCL all-reduce as fallback) ->
alpha Newton update -> ELBO.  Steps cycle through iterations 1..10 of a training run started from
the initial state (the protocol of the reference's published "10 iterations" chart, plots.R:4);
the re-initialisation every 10 steps is outside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = documents processed per second by the whole job with all
inputs resident in HBM; `e2e` = the same metric through the public API (`train(model, iter=1)`) with
host buffers, i.e. including update_buffer! (H2D corpus + parameters) and update_host! (D2H) per step.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 50
CYCLE = 10          # iterations per training run from init (reference chart: 10 iterations)
VITER = 10
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback ("of fallback")


def load_corpus(tm, which):
    c = None
    if which in ("auto", "nsf"):
        c = tm.synth.load_packed("nsf")
        if c is None and which == "nsf":
            raise SystemExit("data/_packed/nsf.npz not found")
    if c is not None:
        return c, "nsf (packed from the reference's datasets/nsf by tools/pack_corpus.py)"
    return tm.synth.nsf_shaped(), "synthetic NSF-shaped corpus (synth.nsf_shaped, seed 1)"


def algorithmic_bytes(nnz, M, V, Kk):
    """per outer iteration, fp32 values + int32 indices."""
    estep = nnz * (8 * Kk + 8) + 12 * Kk * M
    return estep, estep + 12 * Kk * V


class ClockSampler(threading.Thread):
    """Samples SM clock + clock-event reasons through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.mask, self.stop_flag, self.max_mhz, self.ok = index, [], 0, False, None, False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                try:
                    self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:
                    pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(v for k, v in self.REASONS.items() if self.mask & k), "samples": len(self.samples)}


def init_state(tm, model, beta0):
    """Back to the constructor's state (gpuLDA.jl:55-62) with the injected beta."""
    from scipy.special import digamma

    model.alpha = np.ones(K, dtype=np.float32)
    model.beta[...] = beta0
    model.Elogtheta[...] = np.float32(-(np.euler_gamma + digamma(K)))
    model.gamma[...] = 1.0
    model.elbo = 0.0


def run_ours(args):
    import torch

    import topicmodelsvb_b200 as tm

    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun --nproc-per-node %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    reducer = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        reducer = tm.dist.Reducer()
    tm.build()

    full, data_desc = load_corpus(tm, args.data)
    weak = args.scaling == "weak"
    if weak and world > 1:       # every rank holds a full NSF-sized corpus of its own
        shard = full if rank == 0 else tm.synth.nsf_shaped(seed=1 + rank)
        M_total = full.M * world
    else:                        # strong: documents d % world == rank of the one corpus
        shard = full.shard(rank, world) if world > 1 else full
        M_total = full.M
    nnz_total = full.nnz * (world if weak else 1)

    pin = tm._lib.pinned_copy
    shard = shard._replace(N_cumsum=pin(shard.N_cumsum), terms=pin(shard.terms), counts=pin(shard.counts))
    beta0 = np.asfortranarray(tm.synth.init_beta(K, full.V, seed=7).T.astype(np.float32))   # (K, V)

    # one explicit stream for everything: the library's kernels, torch's L2 flush and timing events, NCCL
    work_stream = torch.cuda.Stream()
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream
    model = tm.gpuLDA(tm.Corpus.from_csr(shard), K, reducer=reducer, M_total=M_total, stream=stream)
    model.beta = pin(beta0)
    model.Elogtheta = pin(model.Elogtheta)
    model.gamma = pin(model.gamma)
    vtol = ntol = 1.0 / K**2
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- resident-in-HBM arm: `value` -------------------------------------------------
    model.update_buffer()
    lib, h = tm._lib.load(), model._handle()

    def reinit_device():
        init_state(tm, model, beta0)
        tm._lib.check(lib.tmvb_lda_upload(h, tm._lib.ptr(model.alpha), model.beta.ctypes.data,
                                          model.Elogtheta.ctypes.data, model.gamma.ctypes.data))

    def one_step():
        model.estep(VITER, vtol, want_elbo=True)
        model.update_beta()                  # all-reduce (N > 1) + normalise
        model.update_alpha(1000, ntol)
        return model.update_elbo(0)

    total = args.warmup + args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    estep_ms, sweeps, elbos = [], [], []
    sampler = ClockSampler(local)
    launches0 = None
    for s in range(total):
        if s % CYCLE == 0:
            reinit_device()
        if s == args.warmup:
            barrier()
            launches0 = model.stats().kernel_launches
            sampler.start()
        flush.zero_()                        # evict the previous step's lines from L2
        if s >= args.warmup:
            ev[s - args.warmup][0].record()
        elbo = one_step()
        if s >= args.warmup:
            ev[s - args.warmup][1].record()
            st = model.stats()
            estep_ms.append(st.estep_ms)
            sweeps.append(st.sweeps)
            elbos.append(elbo)
    barrier()
    clocks = sampler.result()
    launches = model.stats().kernel_launches - launches0
    t_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([t_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    t_ms = float(t.item())
    ms_per_step = t_ms / args.steps
    value = M_total * args.steps / (t_ms * 1e-3)

    # ---------------- end-to-end arm through the public API: `e2e` ---------------------------------
    st0 = model.stats()
    e2e_steps = max(3, min(args.steps, CYCLE))
    init_state(tm, model, beta0)
    tm.train(model, iter=1, tol=0.0, viter=VITER, checkelbo=1, printelbo=False)   # warm-up call
    init_state(tm, model, beta0)
    st0 = model.stats()
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        tm.train(model, iter=1, tol=0.0, viter=VITER, checkelbo=1, printelbo=False)
    barrier()
    e2e_s = time.perf_counter() - t0
    st1 = model.stats()
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(te, op=torch.distributed.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e = {"value": M_total * e2e_steps / e2e_s, "unit": "docs/s",
           "h2d_bytes_per_step": (st1.h2d_bytes - st0.h2d_bytes) // e2e_steps,
           "d2h_bytes_per_step": (st1.d2h_bytes - st0.d2h_bytes) // e2e_steps,
           "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
           "call": "train(model, iter=1, checkelbo=1): update_buffer! + E-step + M-step + alpha + ELBO + update_host!"}

    # ---------------- roofline of the dominant kernel (lda_estep_kernel) ---------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    est_bytes_total, iter_bytes_total = algorithmic_bytes(nnz_total, M_total, full.V, K)
    est_ms = float(np.mean(estep_ms))
    achieved = (est_bytes_total / world) / (est_ms * 1e-3) / 1e9          # per GPU
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_lda_estep_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_estep")
    roofline = {"bound": "hbm", "kernel": "lda_estep_reg_kernel / lda_estep_kernel (all length-bucket launches of one E-step)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": est_bytes_total // world, "kernel_ms": est_ms,
                "share_of_step": est_ms / ms_per_step}

    out = {
        "metric": "LDA K=50 NSF: documents/sec over full VI iterations (E-step + M-step + alpha + ELBO)",
        "value": value, "unit": "docs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling if world > 1 else "strong",
        "vs_baseline": None, "dtype": "f32", "data": data_desc,
        "config": {"workload": "gpuLDA K=50 on NSF (128804 docs x 25319 vocab), doc-sharded d % N",
                   "M": int(M_total), "V": int(full.V), "K": K, "nnz": int(nnz_total), "viter": VITER, "vtol": vtol,
                   "step": "one outer VI iteration; steps cycle through iterations 1..%d from the initial state" % CYCLE,
                   "l2": "256 MiB buffer written between timed steps (L2 flushed)", "parallelism": "dp%d" % world,
                   "exchange": ("none (one GPU)" if world == 1 else
                                "fused peer-memory kernel (tmvb_lda_exchange_mstep: reduce-scatter + normalise + all-gather over NVLink)"
                                if getattr(model, "_p2p", False) else "NCCL all-reduce + normalisation kernels")},
        "vi_iterations_per_sec": 1e3 / ms_per_step,
        "estep_docs_per_sec": M_total / (est_ms * 1e-3),
        "sweeps_per_doc": float(np.mean(sweeps)) / (M_total if world > 1 else shard.M),
        "elbo_last": elbos[-1],
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
    }

    # ---------------- CPU baseline: the fp64 oracle port, 1 thread (the reference is single-threaded)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(tm, full, beta0, nthreads=1, docs=24576, iters=2)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


def cpu_baseline(tm, full, beta0, nthreads, docs, iters):
    """Times oracle/lda_oracle.c (port of LDA.jl train!) on the first `docs` documents: `iters` full
    outer iterations (E-step + M-step + alpha + ELBO) from the initial state."""
    import oracle

    sub = tm.synth.take_docs(full, np.arange(min(docs, full.M)))
    st = oracle.LDAState(K, sub.M, sub.V, beta=np.ascontiguousarray(beta0.T))
    t0 = time.perf_counter()
    trace, sw, done = oracle.lda_train(st, sub.N_cumsum, sub.terms, sub.counts, iter=iters, tol=0.0, viter=VITER,
                                       checkelbo=1, nthreads=nthreads)
    dt = time.perf_counter() - t0
    # the initial update_elbo! (LDA.jl:167) is part of train! but not of a steady-state iteration: ~1/(2 iters + 1) of the time
    per_iter = dt / (iters + 0.5)
    return {"value": sub.M / per_iter, "unit": "docs/s", "cores": nthreads, "kind": "port",
            "sample": "first %d NSF documents, %d outer iterations of oracle/lda_oracle.c (fp64 restatement of LDA.jl train!)" % (sub.M, iters),
            "seconds": dt}


def run_reference(args):
    """--impl reference: the reference's CPU train! (its fp64 port -- Julia is not installed) on all host threads."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import oracle
    import topicmodelsvb_b200.synth as synth

    class _TM:  # the oracle arm must not touch the CUDA package beyond the corpus generator
        pass

    tm = _TM()
    tm.synth = synth
    full, data_desc = load_corpus(tm, args.data)
    beta0 = np.asfortranarray(synth.init_beta(K, full.V, seed=7).T.astype(np.float32))
    nthreads = oracle.host_threads()
    docs = 32768
    sub = synth.take_docs(full, np.arange(min(docs, full.M)))
    state = {}

    def reinit():
        state["st"] = oracle.LDAState(K, sub.M, sub.V, beta=np.ascontiguousarray(beta0.T))

    def step():
        oracle.lda_train(state["st"], sub.N_cumsum, sub.terms, sub.counts, iter=1, tol=0.0, viter=VITER,
                         checkelbo=0, nthreads=nthreads)
        oracle.lda_elbo(state["st"], sub.N_cumsum, sub.terms, sub.counts, nthreads=nthreads)

    t_total = 0.0
    for s in range(args.warmup + args.steps):
        if s % CYCLE == 0:
            reinit()
        t0 = time.perf_counter()
        step()
        if s >= args.warmup:
            t_total += time.perf_counter() - t0
    value = sub.M * args.steps / t_total
    out = {
        "impl": "reference",
        "metric": "LDA K=50 NSF: documents/sec over full VI iterations (E-step + M-step + alpha + ELBO)",
        "value": value, "unit": "docs/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": data_desc,
        "config": {"workload": "gpuLDA K=50 on NSF (128804 docs x 25319 vocab), doc-sharded d % N", "K": K, "V": int(full.V),
                   "viter": VITER, "sample_docs": int(sub.M),
                   "step": "one outer VI iteration on a bounded sample; steps cycle through iterations 1..%d from the initial state" % CYCLE},
        "cpu_baseline": {"value": value, "unit": "docs/s", "cores": nthreads, "kind": "port",
                         "sample": "first %d NSF documents per step, oracle/lda_oracle.c (fp64 restatement of LDA.jl train!; "
                                   "Julia is not installed, the reference itself cannot run), OpenMP over documents" % sub.M},
        "e2e": {"value": value, "unit": "docs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--data", default="auto", choices=["auto", "nsf", "synthetic"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
