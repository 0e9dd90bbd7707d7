#!/usr/bin/env python
"""bench.py -- the CAVI hot path of TopicModelsVB.jl on 1..N B200s, one JSON line.

Configurations (BASELINE.json `configs`):
    nsf_lda_k50     gpuLDA  K=50  on NSF (128 804 docs x 25 319 vocab)            configs[1], the headline metric (default)
    citeu_ctm_k30   gpuCTM  K=30  on CiteULike (16 980 docs x 8 000 vocab)        configs[2]
    citeu_ctpf_k30  gpuCTPF K=30  on CiteULike with 5 551 users                   configs[3]
    synth_lda_k200  gpuLDA  K=200 on synthetic 1 M docs x 50 k vocab              configs[4], the HBM-bound scaling sweep

A "step" is ONE outer VI iteration over the whole corpus: fused E-step (all inner sweeps of every document + scatter of
the sufficient statistics) -> M-step (N > 1: one fused peer-memory kernel that reduce-scatters the statistics over
NVLink, normalises and all-gathers; NCCL all-reduce as fallback) -> global updates (alpha Newton / sigma, mu / Gamma
rates) -> ELBO.  Steps cycle through iterations 1..CYCLE of a training run started from the initial state (the protocol
of the reference's published "10 iterations" chart, plots.R:4); the re-initialisation is outside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME] [--also all|none|a,b]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0): the line of `--config`; the lines of the configurations named by `--also` (default at
N = 1: the other three) ride in `"configs": {name: line}` of the same object, each with its own `roofline`,
`cpu_baseline`, `e2e` and `parity`.

`value` = documents per second of the whole job with all inputs resident in HBM; `e2e` = the same metric through the
public API (`train(model, iter=1)`) with host buffers, i.e. including update_buffer! (H2D corpus + parameters) and
update_host! (D2H) in every step; `e2e_iter10` = one `train(model, iter=CYCLE)` call (the reference chart's protocol);
`parity` = max relative ELBO deviation of that call from the committed fp64 oracle trace (tests/golden/bench_traces.json,
made by tools/make_bench_golden.py -- bench.py runs the oracle only in its cpu_baseline / --impl reference legs).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VITER = 10
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback ("of fallback")

CONFIGS = {
    "nsf_lda_k50": dict(model="lda", K=50, corpus="nsf", cycle=10,
                        workload="gpuLDA K=50 on NSF (128804 docs x 25319 vocab), doc-sharded d % N",
                        metric="LDA K=50 NSF: documents/sec over full VI iterations (E-step + M-step + alpha + ELBO)",
                        cpu_docs=24576, cpu_iters=2, ref_docs=None),
    "citeu_ctm_k30": dict(model="ctm", K=30, corpus="citeu", cycle=10,
                          workload="gpuCTM K=30 on CiteULike (16980 docs x 8000 vocab), logistic-normal path, doc-sharded d % N",
                          metric="CTM K=30 CiteULike: documents/sec over full VI iterations (E-step + M-step + sigma/mu + ELBO)",
                          cpu_docs=4096, cpu_iters=2, ref_docs=None),
    "citeu_ctpf_k30": dict(model="ctpf", K=30, corpus="citeu", cycle=10,
                           workload="gpuCTPF K=30 on CiteULike (16980 docs x 8000 vocab x 5551 users), doc-sharded d % N",
                           metric="CTPF K=30 CiteULike: documents/sec over full VI iterations (E-step + M-step + rates + ELBO)",
                           cpu_docs=2048, cpu_iters=2, ref_docs=None),
    "synth_lda_k200": dict(model="lda", K=200, corpus="cfg4", cycle=5,
                           workload="gpuLDA K=200 on synthetic 1M docs x 50k vocab (synth.cfg4_shard), block-cyclic doc shards",
                           metric="LDA K=200 synthetic 1M x 50k: documents/sec over full VI iterations (E-step + M-step + alpha + ELBO)",
                           cpu_docs=4096, cpu_iters=1, ref_docs=8192),
    # the filtered models (fLDA.jl / fCTM.jl; no GPU version in the reference): measured like the others
    "nsf_flda_k50": dict(model="flda", K=50, corpus="nsf", cycle=5,
                         workload="gpufLDA K=50 on NSF (128804 docs x 25319 vocab), filtered LDA (per-token tau, kappa, eta), doc-sharded d % N",
                         metric="fLDA K=50 NSF: documents/sec over full VI iterations (E-step + M-step + alpha/eta + ELBO)",
                         cpu_docs=8192, cpu_iters=1, ref_docs=None),
    "citeu_fctm_k30": dict(model="fctm", K=30, corpus="citeu", cycle=5,
                           workload="gpufCTM K=30 on CiteULike (16980 docs x 8000 vocab), filtered CTM, doc-sharded d % N",
                           metric="fCTM K=30 CiteULike: documents/sec over full VI iterations (E-step + M-step + sigma/mu/kappa + ELBO)",
                           cpu_docs=2048, cpu_iters=1, ref_docs=None),
}
DEFAULT_CONFIG = "nsf_lda_k50"

# What the reference publishes for these workloads (wall-clock bar charts, Apple M1, values in plots.R:4) -- other
# hardware, train! wall time of 10 iterations incl. everything; context only: `vs_baseline` stays null (no number for B200 / this metric)
PUBLISHED = {
    "nsf_lda_k50": {"source": "plots.R:4 (README chart images/gpubar.png), Apple M1, 10 outer iterations of train! on NSF K=50",
                    "reference_cpu_LDA_seconds_per_10_iterations": 444.0, "reference_gpuLDA_opencl_seconds_per_10_iterations": 26.0,
                    "reference_gpuLDA_opencl_docs_per_sec_upper_bound": 49540.0,
                    "compare_with": "e2e_iter10 of this line (one train(iter=10) call incl. update_buffer! / update_host!)"},
}


# ------------------------------------------------------------------------------------------------ corpora --------
def load_corpus(synth, cfg, rank, world, data):
    """(this rank's shard, M_total, nnz_total, V, U, description, golden key or None)."""
    kind = cfg["corpus"]
    if kind == "cfg4":
        M = int(os.environ.get("TMVB_CFG4_M", synth.CFG4_M))
        shard = synth.cfg4_shard(rank, world, M=M)
        return shard, M, None, synth.CFG4_V, 0, "synthetic cfg4 (synth.cfg4_shard: LogNormal lengths, Zipf(1.07) terms, seeds [1, block])", None
    packed = synth.load_packed(kind) if data in ("auto", "packed") else None
    if packed is None and data == "packed":
        raise SystemExit("data/_packed/%s.npz not found" % kind)
    if packed is not None:
        full, desc, golden = packed, "%s (packed from the reference's datasets/%s by tools/pack_corpus.py)" % (kind, kind), True
    else:
        full = synth.nsf_shaped() if kind == "nsf" else synth.citeu_shaped()
        desc, golden = "synthetic %s-shaped corpus (synth.%s_shaped)" % (kind, kind), False
    shard = full.shard(rank, world) if world > 1 else full
    return shard, full.M, full.nnz, full.V, full.U, desc, golden


def algorithmic_bytes(cfg, nnz, M, V, U, nr):
    """bytes one E-step must move (fp32 values, int32 indices), and one whole outer iteration."""
    K = cfg["K"]
    if cfg["model"] == "lda":
        e = nnz * (8 * K + 8) + 12 * K * M
        return e, e + 12 * K * V
    if cfg["model"] == "flda":   # + tau read, tau / tau_old written, kappa weight read per token; the log2 table and kappa in the M-step
        e = nnz * (8 * K + 8 + 16) + 12 * K * M
        return e, e + 16 * K * V + 12 * V
    if cfg["model"] == "ctm":
        e = nnz * (8 * K + 8) + M * 4 * (3 * K + 2)
        return e, e + 12 * K * V + 4 * K * K
    if cfg["model"] == "fctm":
        e = nnz * (8 * K + 8 + 16) + M * 4 * (3 * K + 2)
        return e, e + 16 * K * V + 4 * K * K + 12 * V
    e = nnz * (8 * K + 8) + nr * (8 * K + 8) + 16 * K * M
    return e, e + 12 * K * (V + U)


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


# ------------------------------------------------------------------------------------------------ model arms ------
def filtered_kappa0(V):
    """The injected initial kappa of the filtered configurations (oracle and device read the same float32 vector)."""
    return np.random.default_rng(8).dirichlet(np.ones(V)).astype(np.float32)


class Arm:
    """What the measurement loop needs from one model family: construction with the injected initial state, the
    constructor state again (`reinit`), one resident outer iteration (`step`), the public train call."""

    def __init__(self, tm, cfg, shard, V, U, M_total, reducer, stream):
        self.tm, self.cfg, self.K = tm, cfg, cfg["K"]
        self.shard, self.V, self.U, self.M_total = shard, V, U, M_total
        K, pin = self.K, tm._lib.pinned_copy
        corp = tm.Corpus.from_csr(shard)
        kind = cfg["model"]
        self.vtol = self.ntol = 1.0 / K**2
        if kind == "lda":
            self.init0 = np.asfortranarray(tm.synth.init_beta(K, V, seed=7).T.astype(np.float32))        # (K, V)
            self.model = m = tm.gpuLDA(corp, K, reducer=reducer, M_total=M_total, stream=stream)
            m.beta, m.Elogtheta, m.gamma = pin(self.init0), pin(m.Elogtheta), pin(m.gamma)
        elif kind == "ctm":
            self.init0 = np.asfortranarray(tm.synth.init_beta(K, V, seed=7).T.astype(np.float32))
            self.model = m = tm.gpuCTM(corp, K, reducer=reducer, M_total=M_total, stream=stream)
            m.beta, m.lam, m.vsq = pin(self.init0), pin(m.lam), pin(m.vsq)
        elif kind in ("flda", "fctm"):
            self.init0 = np.asfortranarray(tm.synth.init_beta(K, V, seed=7).T.astype(np.float32))
            self.kappa0 = filtered_kappa0(V)
            cls = tm.gpufLDA if kind == "flda" else tm.gpufCTM
            kw = {}
            if kind == "flda":   # update_eta! divides by sum(model.C) over ALL ranks
                c_local = float(np.asarray(shard.counts, dtype=np.float64).sum())
                kw = dict(C_total=reducer.allreduce_host(c_local) if reducer is not None else c_local)
            self.model = m = cls(corp, K, reducer=reducer, M_total=M_total, stream=stream, **kw)
            m.beta, m.kappa = pin(self.init0), self.kappa0.copy()
            m.tau = pin(m.tau)
            if kind == "flda":
                m.Elogtheta, m.gamma = pin(m.Elogtheta), pin(m.gamma)
            else:
                m.lam, m.vsq = pin(m.lam), pin(m.vsq)
        else:
            self.init0 = np.asfortranarray(tm.synth.init_alef(K, V, seed=7).T.astype(np.float32))
            self.model = m = tm.gpuCTPF(corp, K, reducer=reducer, M_total=M_total, stream=stream)
            m.alef, m.gimel, m.zayin, m.he = pin(self.init0), pin(m.gimel), pin(m.zayin), pin(m.he)

    def reinit_host(self):
        """Back to the constructor's state (gpuLDA.jl:55-62 / gpuCTM.jl:63-72 / gpuCTPF.jl:107-119) with the injected table."""
        from scipy.special import digamma

        m, K, kind = self.model, self.K, self.cfg["model"]
        if kind == "lda":
            m.alpha = np.ones(K, dtype=np.float32)
            m.beta[...] = self.init0
            m.Elogtheta[...] = np.float32(-(np.euler_gamma + digamma(K)))
            m.gamma[...] = 1.0
        elif kind == "ctm":
            m.mu = np.zeros(K, np.float32)
            m.sigma = np.eye(K, dtype=np.float32)
            m.invsigma = np.eye(K, dtype=np.float32)
            m.beta[...] = self.init0
            m.lam[...] = 0.0
            m.vsq[...] = 1.0
            m.logzeta = np.full(m.M, 0.5, np.float32)
        elif kind in ("flda", "fctm"):
            m.eta = 0.5
            m.kappa = self.kappa0.copy()
            m.beta[...] = self.init0          # in place: the fields stay views of page-locked memory
            m.tau[...] = 0.5
            if kind == "flda":
                m.alpha = np.ones(K, dtype=np.float32)
                m.Elogtheta[...] = np.float32(-(np.euler_gamma + digamma(K)))
                m.gamma[...] = 1.0
            else:
                m.mu = np.zeros(K, np.float32)
                m.sigma = np.eye(K, dtype=np.float32)
                m.invsigma = np.eye(K, dtype=np.float32)
                m.lam[...] = 0.0
                m.vsq[...] = 1.0
                m.logzeta = np.full(m.M, 0.5, np.float32)
        else:
            m.alef[...] = self.init0
            m.he[...] = 1.0
            m.gimel[...] = 1.0
            m.zayin[...] = 1.0
            for n in ("bet", "vav", "dalet", "het"):
                setattr(m, n, np.ones(K, np.float32))
        m.elbo = 0.0

    def reinit_device(self):
        self.reinit_host()
        m, lib, P = self.model, self.tm._lib.load(), self.tm._lib.ptr
        if self.cfg["model"] == "lda" and m._resident:   # parameters only: the corpus is already on the device
            self.tm._lib.check(lib.tmvb_lda_upload(m._handle(), P(m.alpha), m.beta.ctypes.data, m.Elogtheta.ctypes.data, m.gamma.ctypes.data))
        else:
            m.update_buffer()

    def step(self):
        m, kind = self.model, self.cfg["model"]
        if kind == "lda":
            if m.can_iterate():              # the whole iteration as one CUDA graph launch (what train() does)
                m.elbo = m.iterate(VITER, self.vtol, 1000, self.ntol, want_elbo=True)
                return m.elbo
            m.estep(VITER, self.vtol, want_elbo=True)
            m.update_beta()                  # all-reduce (N > 1) + normalise
            m.update_alpha(1000, self.ntol)
        elif kind == "ctm":
            m.estep(1000, self.ntol, VITER, self.vtol, want_elbo=True)
            m.mstep()
        elif kind == "flda":
            m.estep(VITER, self.vtol)
            m.mstep(1000, self.ntol)
            return m.update_elbo()
        elif kind == "fctm":
            m.estep(1000, self.ntol, VITER, self.vtol, want_elbo=False)
            m.mstep()
            return m.update_elbo()
        else:
            m.estep(VITER, self.vtol, want_elbo=True)
            m.mstep()
        return m.update_elbo(0)

    def train(self, iters, trace=None):
        self.tm.train(self.model, iter=iters, tol=0.0, viter=VITER, checkelbo=1, printelbo=False, trace=trace)


def measure(tm, torch, args, name, rank, local, world, reducer, work_stream, peak, peak_src):
    """One configuration's JSON line (returned on every rank; only rank 0's cpu_baseline is filled)."""
    cfg = CONFIGS[name]
    K, cycle = cfg["K"], cfg["cycle"]
    shard, M_total, nnz_total, V, U, data_desc, golden_ok = load_corpus(tm.synth, cfg, rank, world, args.data)
    nr_total = 0
    if world > 1 or nnz_total is None:
        t = torch.tensor([shard.nnz, shard.M, int(shard.R_cumsum[-1]) if shard.R_cumsum is not None else 0], dtype=torch.int64, device="cuda")
        if world > 1:
            torch.distributed.all_reduce(t)
        nnz_total, M_sum, nr_total = (int(x) for x in t.tolist())
        assert M_sum == M_total, (M_sum, M_total)
    elif shard.R_cumsum is not None:
        nr_total = int(shard.R_cumsum[-1])
    pin = tm._lib.pinned_copy
    shard = shard._replace(N_cumsum=pin(shard.N_cumsum), terms=pin(shard.terms), counts=pin(shard.counts))
    arm = Arm(tm, cfg, shard, V, U, M_total, reducer, work_stream.cuda_stream)
    model = arm.model
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # ---------------- resident-in-HBM arm: `value` -------------------------------------------------
    steps, warmup = args.steps, args.warmup
    if name == "synth_lda_k200":              # ~100 ms per step: keep the configuration within seconds
        steps, warmup = min(steps, 10), min(warmup, 5)
    model.update_buffer()
    total = warmup + steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    estep_ms, sweeps, elbos = [], [], []
    sampler = ClockSampler(local)
    launches0 = None
    for s in range(total):
        if s % cycle == 0:
            arm.reinit_device()
            if world > 1:
                barrier()   # the re-initialisation is outside the timed region on EVERY rank: nobody's next step waits for a peer's upload
        if s == warmup - 1:
            sampler.start()                  # the NVML thread is up (and has taken its first sample) before the timed region opens
        if s == warmup:
            barrier()
            launches0 = model.stats().kernel_launches
        flush.zero_()                        # evict the previous step's lines from L2
        if s >= warmup:
            ev[s - warmup][0].record()
        elbo = arm.step()
        if s >= warmup:
            ev[s - warmup][1].record()
            st = model.stats()
            estep_ms.append(st.estep_ms)
            sweeps.append(st.sweeps)
            elbos.append(elbo)
    barrier()
    clocks = sampler.result()
    launches = model.stats().kernel_launches - launches0
    per_step = [a.elapsed_time(b) for a, b in ev]
    t_ms = allmax(sum(per_step))
    ms_per_step = t_ms / steps
    value = M_total * steps / (t_ms * 1e-3)

    # ---------------- end-to-end arm through the public API: `e2e` ---------------------------------
    e2e_steps = max(3, min(steps, cycle))
    arm.reinit_host()
    arm.train(1)                             # warm-up call
    arm.reinit_host()
    st0 = model.stats()
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        arm.train(1)
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    st1 = model.stats()
    e2e = {"value": M_total * e2e_steps / e2e_s, "unit": "docs/s",
           "h2d_bytes_per_step": (st1.h2d_bytes - st0.h2d_bytes) // e2e_steps,
           "d2h_bytes_per_step": (st1.d2h_bytes - st0.d2h_bytes) // e2e_steps,
           "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
           "call": "train(model, iter=1, checkelbo=1): update_buffer! + E-step + M-step + global updates + ELBO + update_host!"}

    # ---------------- one train(iter=CYCLE) call: the reference chart's protocol + ELBO parity ------
    arm.reinit_host()
    trace = []
    st0 = model.stats()
    barrier()
    t0 = time.perf_counter()
    arm.train(cycle, trace)
    barrier()
    it_s = allmax(time.perf_counter() - t0)
    st1 = model.stats()
    e2e_iter = {"value": M_total * cycle / it_s, "unit": "docs/s", "ms_per_iteration": 1e3 * it_s / cycle, "iterations": cycle,
                "h2d_bytes_per_call": st1.h2d_bytes - st0.h2d_bytes, "d2h_bytes_per_call": st1.d2h_bytes - st0.d2h_bytes,
                "call": "train(model, iter=%d, checkelbo=1): one update_buffer!, %d outer iterations, one update_host!" % (cycle, cycle)}
    parity = parity_block(name, cfg, trace, golden_ok, tm, torch, rank, world, reducer, work_stream)

    # ---------------- roofline of the dominant kernel (the E-step launches) -------------------------
    est_bytes_total, iter_bytes_total = algorithmic_bytes(cfg, nnz_total, M_total, V, U, nr_total)
    est_ms = float(np.mean(estep_ms))
    achieved = (est_bytes_total / world) / (est_ms * 1e-3) / 1e9          # per GPU
    traffic, tsrc = None, None
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tpath) and world == 1:   # the capture is of the whole corpus on one GPU
        tj = json.load(open(tpath)).get(name)
        if tj:
            traffic, tsrc = tj.get("dram_bytes_per_estep"), tj.get("source")
    kernel_names = {"lda": "lda_estep_hyb_kernel / lda_estep_kernel", "ctm": "ctm_estep_kernel", "ctpf": "ctpf_estep_kernel",
                    "flda": "flda_estep_reg_kernel / flda_estep_kernel", "fctm": "fctm_estep_kernel"}
    # what ncu names as the limiter of each kernel (profiles/r2_*_ncu_full_summary.txt): none of these working sets is HBM-resident at its
    # configuration, so `frac` (algorithmic bytes against the HBM peak, the contract's roofline) is a lower bound on distance, not the limiter
    limiter = {"lda": "warp issue / fixed-latency FFMA2 chains at register-limited occupancy (8-12 warps per SM); DRAM traffic is 4 % of the algorithmic bytes (table and slab L2- / register-resident)",
               "ctm": "instruction issue in the per-document Newton iterations (register-resident Cholesky + triangular solves: ~60 % of the instructions)",
               "ctpf": "instruction issue / latency at 12 % occupancy (two token passes over two tiles per sweep, digamma per topic)",
               "flda": "warp issue (44-56 %) and the XU pipe (40-47 %: one MUFU.EX2 per token, topic and sweep; its floor is 1.7 ms) at 168 registers / 12 resident warps per SM",
               "fctm": "instruction issue in the Newton iterations, as ctm, plus one exponential per token and topic"}[cfg["model"]]
    roofline = {"bound": "hbm", "limiter": limiter, "kernel": kernel_names[cfg["model"]] + " (all length-bucket launches of one E-step)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": est_bytes_total // world, "kernel_ms": est_ms,
                "share_of_step": est_ms / ms_per_step}

    out = {
        "metric": cfg["metric"], "value": value, "unit": "docs/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "ms_per_step_min_med_max": [min(per_step), statistics.median(per_step), max(per_step)],
        "ms_per_step_rank0": [round(x, 4) for x in per_step],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": data_desc,
        "config": {"workload": cfg["workload"], "name": name, "M": int(M_total), "V": int(V), "K": K, "nnz": int(nnz_total), "viter": VITER,
                   "vtol": arm.vtol, "step": "one outer VI iteration; steps cycle through iterations 1..%d from the initial state" % cycle,
                   "l2": "256 MiB buffer written between timed steps (L2 flushed)", "parallelism": "dp%d" % world,
                   "exchange": ("none (one GPU)" if world == 1 else
                                "fused peer-memory kernel (reduce-scatter + normalise + all-gather over NVLink)"
                                if getattr(model, "_p2p", False) else "NCCL all-reduce + normalisation kernels")},
        "published_context": PUBLISHED.get(name),
        "vi_iterations_per_sec": 1e3 / ms_per_step,
        "estep_docs_per_sec": M_total / (est_ms * 1e-3),
        "sweeps_per_doc": float(np.mean(sweeps)) / (M_total if world > 1 else shard.M),
        "elbo_last": elbos[-1],
        "e2e": e2e, "e2e_iter%d" % cycle: e2e_iter, "parity": parity,
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
    }
    if U:
        out["config"]["U"] = int(U)
    # ---------------- CPU baseline: the fp64 oracle port, 1 thread (the reference is single-threaded)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(tm.synth, cfg, shard, arm.init0, nthreads=1, docs=cfg["cpu_docs"], iters=cfg["cpu_iters"])
    model.close()
    del arm, model, flush
    torch.cuda.empty_cache()
    return out


def parity_block(name, cfg, trace, golden_ok, tm, torch, rank, world, reducer, work_stream):
    """ELBO after every outer iteration of the train(iter=CYCLE) call vs the committed oracle trace."""
    gpath = os.path.join(ROOT, "tests", "golden", "bench_traces.json")
    if not os.path.exists(gpath):
        return {"max_rel_vs_oracle": None, "why": "tests/golden/bench_traces.json missing"}
    gold = json.load(open(gpath)).get(name)
    if gold is None:
        return {"max_rel_vs_oracle": None, "why": "no oracle trace for this configuration"}
    if cfg["corpus"] == "cfg4":
        # the oracle trace covers the first two blocks of the corpus (the fp64 oracle needs ~10 min per iteration on the
        # full million): train those 25 000 documents, sharded d % N like any other corpus, and compare
        sub = tm.synth.cfg4_shard(0, 1, M=int(gold["M"]))
        if world > 1:
            sub = sub.shard(rank, world)
        K = cfg["K"]
        m = tm.gpuLDA(tm.Corpus.from_csr(sub), K, reducer=reducer, M_total=int(gold["M"]), stream=work_stream.cuda_stream)
        m.beta = np.asfortranarray(tm.synth.init_beta(K, sub.V, seed=7).T.astype(np.float32))
        trace = []
        tm.train(m, iter=int(gold["iter"]), tol=0.0, viter=VITER, checkelbo=1, printelbo=False, trace=trace)
        m.close()
        what = "first %d documents of the corpus, %d iterations" % (gold["M"], gold["iter"])
    elif not golden_ok:
        return {"max_rel_vs_oracle": None, "why": "packed corpus not on this box (synthetic stand-in): the oracle trace does not apply"}
    else:
        what = "full corpus, %d iterations" % gold["iter"]
    ref = np.array(gold["elbo"], dtype=np.float64)
    got = np.array(trace, dtype=np.float64)
    n = min(len(ref), len(got))
    rel = np.abs(got[:n] - ref[:n]) / np.abs(ref[:n])
    return {"max_rel_vs_oracle": float(rel.max()), "iters": int(n - 1), "what": what, "tolerance": 1e-4,
            "elbo_final": float(got[n - 1]), "oracle_elbo_final": float(ref[n - 1]),
            "source": "tests/golden/bench_traces.json (fp64 oracle port of the CPU train!; parity with the reference itself is unpinned)"}


# ------------------------------------------------------------------------------------------------ CPU arms --------
def oracle_runner(synth, cfg, sub, init0, nthreads):
    """(reinit, step) closures over the fp64 oracle: one step = one outer iteration incl. the ELBO (checkelbo=1)."""
    import oracle

    K, kind = cfg["K"], cfg["model"]
    state = {}
    table = np.ascontiguousarray(init0.T)

    def reinit():
        if kind == "lda":
            state["st"] = oracle.LDAState(K, sub.M, sub.V, beta=table)
        elif kind == "ctm":
            state["st"] = oracle.CTMState(K, sub.M, sub.V, table)
        elif kind == "flda":
            state["st"] = oracle.FLDAState(K, sub.M, sub.V, len(sub.terms), table, filtered_kappa0(sub.V))
        elif kind == "fctm":
            state["st"] = oracle.FCTMState(K, sub.M, sub.V, len(sub.terms), table, filtered_kappa0(sub.V))
        else:
            state["st"] = oracle.CTPFState(K, sub.M, sub.V, sub.U, table)

    def step():
        st = state["st"]
        if kind == "lda":
            oracle.lda_train(st, sub.N_cumsum, sub.terms, sub.counts, iter=1, tol=0.0, viter=VITER, checkelbo=0, nthreads=nthreads)
            oracle.lda_elbo(st, sub.N_cumsum, sub.terms, sub.counts, nthreads=nthreads)
        elif kind == "ctm":
            # checkelbo=1 evaluates the ELBO before and after the iteration: count one of the two
            oracle.ctm_train(st, sub.N_cumsum, sub.terms, sub.counts, iter=1, tol=0.0, viter=VITER, checkelbo=1, nthreads=nthreads)
        elif kind == "flda":
            oracle.flda_train(st, sub.N_cumsum, sub.terms, sub.counts, iter=1, tol=0.0, viter=VITER, checkelbo=1, nthreads=nthreads)
        elif kind == "fctm":
            oracle.fctm_train(st, sub.N_cumsum, sub.terms, sub.counts, iter=1, tol=0.0, viter=VITER, checkelbo=1, nthreads=nthreads)
        else:
            oracle.ctpf_train(st, sub, iter=1, tol=0.0, viter=VITER, checkelbo=1, nthreads=nthreads)

    return reinit, step


def cpu_baseline(synth, cfg, full, init0, nthreads, docs, iters):
    """Times the oracle port of the CPU train! on the first `docs` documents: `iters` outer iterations from the initial state."""
    sub = synth.take_docs(full, np.arange(min(docs, full.M)))
    reinit, step = oracle_runner(synth, cfg, sub, init0, nthreads)
    reinit()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    dt = time.perf_counter() - t0
    src = {"lda": "oracle/lda_oracle.c (fp64 restatement of LDA.jl train!)", "ctm": "oracle/ctm_oracle.c (fp64 restatement of CTM.jl train!)",
           "flda": "oracle/flda_oracle.c (fp64 restatement of fLDA.jl train!)", "fctm": "oracle/fctm_oracle.c (fp64 restatement of fCTM.jl train!)",
           "ctpf": "oracle/ctpf_oracle.c (fp64 restatement of CTPF.jl train!, long-form ELBO)"}[cfg["model"]]
    return {"value": sub.M * iters / dt, "unit": "docs/s", "cores": nthreads, "kind": "port",
            "sample": "first %d documents, %d outer iterations of %s" % (sub.M, iters, src), "seconds": dt}


def run_reference(args):
    """--impl reference: the reference's CPU train! (its fp64 port -- Julia is not installed) on all host threads."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import oracle
    import topicmodelsvb_b200.synth as synth

    name = args.config
    cfg = CONFIGS[name]
    K, cycle = cfg["K"], cfg["cycle"]
    if cfg["corpus"] == "cfg4":   # a bounded sample: the first blocks of the corpus (the fp64 port needs ~10 min per iteration on the million)
        full, M_total, V, U = synth.cfg4_shard(0, 1, M=cfg["ref_docs"]), synth.CFG4_M, synth.CFG4_V, 0
        data_desc = "synthetic cfg4 (synth.cfg4_shard), first %d documents" % cfg["ref_docs"]
    else:
        full, M_total, _, V, U, data_desc, _ = load_corpus(synth, cfg, 0, 1, args.data)
    init = synth.init_alef if cfg["model"] == "ctpf" else synth.init_beta
    init0 = np.asfortranarray(init(K, V, seed=7).T.astype(np.float32))
    nthreads = oracle.host_threads()
    docs = cfg["ref_docs"] or full.M
    sub = full if docs >= full.M else synth.take_docs(full, np.arange(docs))
    reinit, step = oracle_runner(synth, cfg, sub, init0, nthreads)
    t_total, per = 0.0, []
    for s in range(args.warmup + args.steps):
        if s % cycle == 0:
            reinit()
        t0 = time.perf_counter()
        step()
        if s >= args.warmup:
            per.append(time.perf_counter() - t0)
    t_total = sum(per)
    value = sub.M * args.steps / t_total
    sample = ("the full corpus (%d documents) per step" % sub.M) if sub.M == M_total else ("first %d documents per step" % sub.M)
    out = {
        "impl": "reference", "metric": cfg["metric"],
        "value": value, "unit": "docs/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": data_desc,
        "config": {"workload": cfg["workload"], "name": name, "K": K, "V": int(V), "M": int(sub.M), "viter": VITER, "sample_docs": int(sub.M),
                   "step": "one outer VI iteration (E-step + M-step + global updates + ELBO); steps cycle through iterations 1..%d from the initial state" % cycle},
        "cpu_baseline": {"value": value, "unit": "docs/s", "cores": nthreads, "kind": "port",
                         "sample": sample + ", oracle/%s_oracle.c (fp64 restatement of the CPU train!; Julia is not installed, the "
                                   "reference itself cannot run), OpenMP over documents" % cfg["model"]},
        "e2e": {"value": value, "unit": "docs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def run_ours(args):
    import torch

    import topicmodelsvb_b200 as tm

    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("--gpus %d needs torchrun --nproc-per-node %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    reducer = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        reducer = tm.dist.Reducer()
    tm.build()
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"

    # one explicit stream for everything: the library's kernels, torch's L2 flush and timing events, NCCL
    work_stream = torch.cuda.Stream()
    torch.cuda.set_stream(work_stream)

    out = measure(tm, torch, args, args.config, rank, local, world, reducer, work_stream, peak, peak_src)
    also = args.also
    if also == "auto":
        # the other BASELINE configurations ride in the same line; with several GPUs only the LDA scaling sweep does
        # (gpuCTM / gpuCTPF shard through the NCCL path of their host mirrors: measured with --config on its own)
        also = "all" if world == 1 else ("synth_lda_k200" if args.config == DEFAULT_CONFIG else "none")
    names = [n for n in CONFIGS if n != args.config] if also == "all" else [n for n in also.split(",") if n and n != "none"]
    if names:
        out["configs"] = {}
    for n in names:
        # a failure must fail on every rank alike (the ranks meet in collectives): no per-rank recovery
        t0 = time.perf_counter()
        try:
            line = measure(tm, torch, args, n, rank, local, world, reducer, work_stream, peak, peak_src)
        except Exception as e:   # a secondary configuration must not take the headline line with it (one process: nobody waits in a collective)
            if world > 1:
                raise
            line = {"error": repr(e)}
        line["bench_seconds"] = time.perf_counter() - t0
        out["configs"][n] = line
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=DEFAULT_CONFIG, choices=sorted(CONFIGS))
    ap.add_argument("--also", default="auto", help="auto | all | none | comma-separated configuration names")
    ap.add_argument("--data", default="auto", choices=["auto", "packed", "synthetic"])
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
