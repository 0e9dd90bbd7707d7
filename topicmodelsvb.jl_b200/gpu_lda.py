"""gpuLDA -- host mirror of the reference's ``gpuLDA`` model and its ``train!`` (src/gpuLDA.jl),
driving libtmvb.so through the C ABI (include/tmvb.h).  Julia is not installed in this image, so this
Python layer stands where the Julia shim of INTEGRATION.md would: same struct fields, same keyword
arguments, same argument checks and error kinds, same order of operations.

Semantics follow the CPU model (src/LDA.jl) where the two reference implementations differ
(per-document stopping rule, lagged-phi ELBO) -- BASELINE.json pins parity to the CPU LDA.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import numpy as np
from scipy.special import digamma

from . import _lib
from .corpus import Corpus, check_corp
from .dist import Reducer


def _fmat(x, K, n, name):
    a = np.asarray(x, dtype=np.float32)
    if a.shape != (K, n):
        raise _lib.TopicModelError("%s must be of size (%d, %d)." % (name, K, n))
    return np.asfortranarray(a)


class gpuLDA:
    """GPU accelerated latent Dirichlet allocation model (gpuLDA.jl:6-85).

    Dense matrices are (K, V) / (K, M) Fortran-ordered float32 arrays -- byte-identical to the
    reference's column-major Float32 matrices, so ``beta[i, j]`` and ``Elogtheta[:, d]`` read as in Julia.
    ``phi`` is materialised lazily (``model.phi``) instead of living in HBM: the fused E-step never
    writes it.

    ``reducer``: a ``dist.Reducer`` when this model holds one shard (documents d % world == rank) of
    a corpus trained on several GPUs; ``M_total`` is then the corpus-wide document count.
    """

    def __init__(self, corp: Corpus, K: int, seed: Optional[int] = None, device: int = -1,
                 reducer: Optional[Reducer] = None, M_total: Optional[int] = None, stream: Optional[int] = None):
        check_corp(corp)
        if not (isinstance(K, (int, np.integer)) and K > 0):
            raise ValueError("number of topics must be a positive integer.")  # gpuLDA.jl:47
        M, V, _ = corp.size()
        corp = corp.copy()      # the reference stores copy(corp): later edits of the caller's corpus do not reach the model
        flat = corp.flat()
        self.K, self.M, self.V = int(K), int(M), int(V)
        self.N = corp.lengths()             # read-only; rebind (not mutate) to change it
        cs = np.concatenate([[0], np.cumsum(flat.counts)]).astype(np.int64)
        self.C = cs[flat.N_cumsum[1:]] - cs[flat.N_cumsum[:-1]]
        self.corp = corp
        self.topics = [np.arange(1, V + 1) for _ in range(K)]
        rng = np.random.default_rng(seed)
        self.alpha = np.ones(K, dtype=np.float32)                                        # gpuLDA.jl:55
        g = rng.standard_exponential(size=(K, V)) if V else np.zeros((K, 0))
        self.beta = np.asfortranarray((g / g.sum(axis=1, keepdims=True)).astype(np.float32)) if V else np.zeros((K, 0), np.float32, order="F")
        e0 = np.float32(-(np.euler_gamma + digamma(K)))                                  # gpuLDA.jl:57
        self.Elogtheta = np.full((K, M), e0, dtype=np.float32, order="F")
        self.Elogtheta_sum = self.Elogtheta.sum(axis=1, dtype=np.float64)
        self.gamma = np.ones((K, M), dtype=np.float32, order="F")                        # gpuLDA.jl:60
        self._beta_old = self.beta.copy(order="F")
        self._Elogtheta_old = self.Elogtheta.copy(order="F")
        self._old_on_device = False   # beta_old / Elogtheta_old are fetched from the device on first access
        self.elbo = 0.0
        self.sweeps = 0
        self.reducer = reducer
        self.M_total = int(M_total) if M_total is not None else self.M
        self._device = device
        self._stream = stream  # caller-owned cudaStream_t (int) or None
        self._h = None
        self._resident = False
        self._pinned = None
        self._corpus_on_device = None   # the Corpus object whose flattening the handle holds

    # ------------------------------------------------------------------ device plumbing ------
    def _handle(self):
        if self._h is None:
            lib = _lib.load()
            h = C.c_void_p()
            stream = self._stream if self._stream is not None else (self.reducer.stream_ptr() if self.reducer is not None else None)
            _lib.check(lib.tmvb_lda_create(C.byref(h), self.K, self.M, self.V, self._device, stream))
            self._h = h
            self._p2p = False
            if self.reducer is not None and self.reducer.world > 1 and os.environ.get("TMVB_P2P", "1") != "0" and self.V > 0:
                # peer-memory exchange: one fused reduce-scatter + normalise + all-gather kernel per iteration instead
                # of NCCL all-reduces + normalisation kernels (falls back to NCCL when the peers cannot be mapped)
                self._p2p = self.reducer.connect_peers(
                    lambda buf, n: _lib.check(lib.tmvb_lda_comm_export(h, buf, n)),
                    lambda rank, world, blobs, n: _lib.check(lib.tmvb_lda_comm_connect(h, rank, world, blobs, n)),
                    _lib.COMM_BLOB_BYTES)
        return self._h

    def close(self):
        if self._h is not None:
            _lib.load().tmvb_lda_destroy(self._h)
            self._h = None
            self._resident = False
            self._corpus_on_device = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update_buffer(self):
        """update_buffer!(model::gpuLDA) (modelutils.jl:370-397): flatten, upload corpus and parameters."""
        lib, h = _lib.load(), self._handle()
        f = self.corp.flat()
        # A corpus wrapped from a flattened CSR (Corpus.from_csr / readcorp) is immutable, so its device copy stays valid for the
        # life of the handle: the reference re-flattens and re-uploads the corpus on every update_buffer! (modelutils.jl:371-388)
        # because its documents are mutable -- a Corpus built from Document objects still takes that path here.
        if not (self.corp.docs is None and self._corpus_on_device is self.corp):
            # the flattened corpus is cached on the host as page-locked Int32 (built once per Corpus): half the upload of the
            # Int64 vectors update_buffer! rebuilds on every call (modelutils.jl:371-373)
            t32, c32 = (None, None) if os.environ.get("TMVB_CORPUS64") == "1" else self.corp.flat32()
            if t32 is not None:
                _lib.check(lib.tmvb_lda_set_corpus32(h, _lib.ptr(f.N_cumsum), _lib.ptr(t32), _lib.ptr(c32)))
            else:
                _lib.check(lib.tmvb_lda_set_corpus(h, _lib.ptr(f.N_cumsum), _lib.ptr(f.terms), _lib.ptr(f.counts)))
            self._corpus_on_device = self.corp
        self.alpha = np.ascontiguousarray(self.alpha, dtype=np.float32)
        self.beta = _fmat(self.beta, self.K, self.V, "beta")
        self.Elogtheta = _fmat(self.Elogtheta, self.K, self.M, "Elogtheta")
        self.gamma = _fmat(self.gamma, self.K, self.M, "gamma")
        _lib.check(lib.tmvb_lda_upload(h, _lib.ptr(self.alpha), self.beta.ctypes.data, self.Elogtheta.ctypes.data,
                                       self.gamma.ctypes.data))
        self._resident = True

    def update_host(self):
        """update_host!(model::gpuLDA) (modelutils.jl:501-516) minus phi (see ``phi``)."""
        if not self._resident:
            return
        lib, h = _lib.load(), self._handle()
        # fresh arrays, as the reference's update_host! rebinds the fields (never write through to caller arrays)
        # page-locked staging arrays (D2H at PCIe speed), allocated once per model: the fields are live
        # views of these buffers and are overwritten by the next update_host! -- copy to keep a snapshot
        pb = self._staging()
        self.alpha = np.empty(self.K, np.float32)
        self.beta, self.Elogtheta, self.gamma = pb["beta"], pb["Elogtheta"], pb["gamma"]
        _lib.check(lib.tmvb_lda_download(h, _lib.ptr(self.alpha), self.beta.ctypes.data, self.Elogtheta.ctypes.data,
                                         self.gamma.ctypes.data))
        self._old_on_device = True
        es = np.zeros(self.K)
        _lib.check(lib.tmvb_lda_get_elogtheta_sum(h, _lib.ptr(es)))
        self.Elogtheta_sum = es

    def _staging(self):
        if self._pinned is None:
            pe = _lib.pinned_empty
            self._pinned = dict(beta=pe((self.K, self.V), np.float32, order="F"), beta_old=pe((self.K, self.V), np.float32, order="F"),
                                Elogtheta=pe((self.K, self.M), np.float32, order="F"),
                                Elogtheta_old=pe((self.K, self.M), np.float32, order="F"),
                                gamma=pe((self.K, self.M), np.float32, order="F"), topics=pe((self.K, self.V), np.int32))
        return self._pinned

    def arm_host_mirror(self):
        """The Elogtheta / gamma half of update_host! (modelutils.jl:507-512) rides with the next E-step: its kernels write each
        document's final rows into the page-locked arrays update_host! hands out, while the other documents are still swept.
        Called by train! before the iteration that is certainly its last (k == iter); a run that stops earlier on `tol`
        downloads as before.  One GPU only (TMVB_HOST_MIRROR=0 switches it off)."""
        if os.environ.get("TMVB_HOST_MIRROR", "1") == "0" or not self.M or (self.reducer is not None and self.reducer.world > 1):
            return
        # (after an earlier train! call model.Elogtheta / model.gamma ARE these arrays: update_buffer!'s copy out of them is
        # ordered before the E-step on the handle's stream, and update_host! rebinds the fields to them anyway)
        pb = self._staging()
        _lib.check(_lib.load().tmvb_lda_arm_host_mirror(self._handle(), pb["Elogtheta"].ctypes.data, pb["gamma"].ctypes.data))

    def _fetch_old(self):
        """beta_old (LDA.jl:122) / Elogtheta_old (LDA.jl:137) live on the device between update_host! calls; the reference's
        update_host! (modelutils.jl:501-516) does not transfer them, so they cross the bus only when somebody looks."""
        if self._old_on_device and self._h is not None:
            pb = self._pinned
            _lib.check(_lib.load().tmvb_lda_download_old(self._handle(), pb["beta_old"].ctypes.data, pb["Elogtheta_old"].ctypes.data))
            self._beta_old, self._Elogtheta_old = pb["beta_old"], pb["Elogtheta_old"]
            self._old_on_device = False

    @property
    def beta_old(self):
        self._fetch_old()
        return self._beta_old

    @beta_old.setter
    def beta_old(self, v):
        self._fetch_old()
        self._beta_old = v

    @property
    def Elogtheta_old(self):
        self._fetch_old()
        return self._Elogtheta_old

    @Elogtheta_old.setter
    def Elogtheta_old(self, v):
        self._fetch_old()
        self._Elogtheta_old = v

    def update_topics(self):
        """model.topics = [reverse(sortperm(vec(beta[i,:]))) for i in 1:K] (gpuLDA.jl:374), ranked on the device."""
        if not self.V:
            return
        if not self._resident:
            self.topics = [np.argsort(self.beta[i, :], kind="stable")[::-1] + 1 for i in range(self.K)]
            return
        if self._pinned is None:
            self.update_host()
        t = self._pinned["topics"]
        _lib.check(_lib.load().tmvb_lda_topics(self._handle(), t.ctypes.data))
        self.topics = list(t)

    @property
    def phi(self):
        """phi[d] (K x N_d) for every document: the last phi of the E-step (LDA.jl:87-88).  K*sum(N) floats."""
        f = self.corp.flat()
        if not self._resident:
            self.update_buffer()
        out = np.zeros((f.nnz, self.K), dtype=np.float32)
        _lib.check(_lib.load().tmvb_lda_materialize_phi(self._handle(), _lib.ptr(out)))
        return [out[f.N_cumsum[d]:f.N_cumsum[d + 1]].T for d in range(self.M)]

    def stats(self) -> _lib.TmvbStats:
        st = _lib.TmvbStats()
        _lib.check(_lib.load().tmvb_lda_get_stats(self._handle(), C.byref(st)))
        return st

    # ------------------------------------------------------------------ the train! steps -----
    def estep(self, viter: int, vtol: float, want_elbo: bool = True):
        """The folded inner loop: update_phi!/update_gamma!/update_Elogtheta! (gpuLDA.jl:356-364) + scatter."""
        _lib.check(_lib.load().tmvb_lda_estep(self._handle(), int(viter), float(vtol), int(bool(want_elbo))))

    def _reduce(self):
        if self.reducer is None:
            return
        lib, h = _lib.load(), self._handle()
        sp, sn, mp, mn = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_int64()
        _lib.check(lib.tmvb_lda_reduce_buffers(h, C.byref(sp), C.byref(sn), C.byref(mp), C.byref(mn)))
        dev = self.reducer.torch.cuda.current_device()
        self.reducer.allreduce_device([(sp.value, sn.value, "<f4"), (mp.value, mn.value, "<f8")], dev)

    def update_beta(self):
        """update_beta!(model::gpuLDA) (gpuLDA.jl:201-204): [all-reduce the statistics,] normalise."""
        h = self._handle()
        if getattr(self, "_p2p", False):
            _lib.check(_lib.load().tmvb_lda_exchange_mstep(h))
            return
        self._reduce()
        _lib.check(_lib.load().tmvb_lda_mstep(h))

    def update_alpha(self, niter: int, ntol: float):
        """update_alpha!(model::gpuLDA, niter, ntol) (gpuLDA.jl:132-154), fp64 as in LDA.jl:97-118."""
        # asynchronous: the Newton iteration runs in an fp64 device kernel; model.alpha is refreshed by update_host!
        _lib.check(_lib.load().tmvb_lda_update_alpha(self._handle(), self.M_total, int(niter), float(ntol), None))

    def can_iterate(self) -> bool:
        """The fused outer iteration (tmvb_lda_iterate) applies on one GPU and with mapped peers; a run that sums the statistics
        through torch.distributed (NCCL / gloo fallback) keeps the separate steps."""
        self._handle()
        return os.environ.get("TMVB_UNFUSED") != "1" and (self.reducer is None or self.reducer.world == 1 or getattr(self, "_p2p", False))

    def iterate(self, viter: int, vtol: float, niter: int, ntol: float, want_elbo: bool = True):
        """One pass of the `for k in 1:iter` body of train! (gpuLDA.jl:355-371) as one CUDA graph launch: E-step, update_beta!,
        update_alpha! and -- with want_elbo -- the ELBO check_elbo! reads (returned; model.elbo is NOT touched here)."""
        e = C.c_double()
        _lib.check(_lib.load().tmvb_lda_iterate(self._handle(), int(viter), float(vtol), int(bool(want_elbo)), self.M_total, int(niter),
                                                float(ntol), C.byref(e)))
        return e.value if want_elbo else None

    def update_elbo(self, mode: int = 0) -> float:
        """update_elbo!(model) (gpuLDA.jl:121-128) from device-side partials; no phi transfer."""
        docs, glob = C.c_double(), C.c_double()
        _lib.check(_lib.load().tmvb_lda_elbo(self._handle(), mode, self.M_total, C.byref(docs), C.byref(glob)))
        d = docs.value
        if mode == 1 and self.reducer is not None:
            d = self.reducer.allreduce_host(d)
        self.elbo = d + glob.value
        return self.elbo


def check_model(model: gpuLDA) -> None:
    """check_model(model::gpuLDA) (modelutils.jl:255-279) -- the invariants that do not need phi."""
    E = _lib.TopicModelError
    K, M, V = model.K, model.M, model.V
    f = model.corp.flat()
    if M != len(model.corp):
        raise E("M must equal the number of documents in the corpus.")
    L = model.corp.lengths()
    if model.N is not L and not np.array_equal(model.N, L):
        raise E("N must contain document lengths.")
    a = np.asarray(model.alpha)
    if a.shape != (K,):
        raise E("alpha must be of length K.")
    if not np.all(np.isfinite(a)):
        raise E("alpha must be finite.")
    if not np.all(a > 0):
        raise E("alpha must be positive.")
    # the element-wise invariants (beta right-stochastic/non-negative, Elogtheta finite and <= 0, gamma
    # finite and > 0; modelutils.jl:264-273) are evaluated on the device copy during update_buffer!
    if np.shape(model.beta) != (K, V):
        raise E("beta must be of size (K, V).")
    # isstochastic(beta, dims=2) (modelutils.jl:268) is evaluated on the device copy during update_buffer! (row sums in fp64 by
    # shard_check_stochastic): the same TopicModelError, without a 5 MB host pass per train! call
    if np.shape(model.Elogtheta) != (K, M):
        raise E("Elogtheta must contain M vectors of length K.")
    if np.shape(model.gamma) != (K, M):
        raise E("gamma must contain M vectors of length K.")
    if not math.isfinite(model.elbo):
        raise E("elbo must be finite.")


def check_elbo(model, checkelbo, printelbo: bool, k: int, tol: float) -> bool:
    """check_elbo!(model, checkelbo, printelbo, k, tol) (modelutils.jl:574-585).  The reference first
    copies the whole model (incl. K x sumN phi) to the host; here the ELBO is assembled from
    device-side partials, so nothing but O(K) doubles crosses the bus."""
    if checkelbo != math.inf and k % checkelbo == 0:
        old = model.elbo
        delta = model.update_elbo(0) - old
        if printelbo:
            print("%d ∆elbo: %.3f" % (k, delta))
        if delta < tol:
            return True
    return False


def train(model: gpuLDA, iter: int = 150, tol: float = 1.0, niter: int = 1000, ntol: Optional[float] = None,
          viter: int = 10, vtol: Optional[float] = None, checkelbo=1, printelbo: bool = True, trace: Optional[list] = None):
    """train!(model::gpuLDA; iter, tol, niter, ntol, viter, vtol, checkelbo, printelbo) (gpuLDA.jl:347-376).

    ``trace`` (optional list) receives the ELBO after every checked iteration (slot 0 = initial).
    """
    K = model.K
    ntol = 1.0 / K**2 if ntol is None else ntol
    vtol = 1.0 / K**2 if vtol is None else vtol
    check_model(model)
    if not all(t >= 0 for t in (tol, ntol, vtol)):
        raise ValueError("tolerance parameters must be nonnegative.")
    if not all(t >= 0 for t in (iter, niter, viter)):
        raise ValueError("iteration parameters must be nonnegative.")
    if not ((isinstance(checkelbo, (int, np.integer)) and checkelbo > 0) or checkelbo == math.inf):
        raise ValueError("checkelbo parameter must be a positive integer or Inf.")
    if iter > 0 and viter < 1:
        # the reference accepts viter = 0 and then scatters whatever phi the previous call left behind (CPU: a scratch matrix of
        # another document's shape); the fused E-step has no stored phi, so the degenerate case is refused up front
        raise ValueError("viter must be at least 1 (the fused E-step does not keep a stale phi to scatter).")
    if model.corp.flat().nnz == 0 and (model.reducer is None):
        iter = 0                                               # gpuLDA.jl:352
    else:
        model.update_buffer()
    check = checkelbo != math.inf
    synced = False
    if check and checkelbo <= iter:
        model.update_elbo(1)                                   # gpuLDA.jl:353
        synced = model.reducer is not None                     # the all-reduce of the initial ELBO: every rank has uploaded
        if trace is not None:
            trace.append(model.elbo)

    fused = iter > 0 and model.can_iterate()
    if model.reducer is not None and model.reducer.world > 1 and iter > 0 and not synced:
        model.reducer.barrier()    # every rank has uploaded before the first exchange (the device barriers only wait so long)
    for k in range(1, iter + 1):
        want = check and (k % checkelbo == 0)
        if k == iter:
            model.arm_host_mirror()                                # update_host! of gamma / Elogtheta inside this E-step
        if fused:
            # gpuLDA.jl:356-368 as one graph launch; check_elbo! (modelutils.jl:574-585) on the value it returns
            new_elbo = model.iterate(viter, vtol, niter, ntol, want_elbo=want)
            stop = False
            if want:
                delta, model.elbo = new_elbo - model.elbo, new_elbo
                if printelbo:
                    print("%d ∆elbo: %.3f" % (k, delta))
                stop = delta < tol
        else:
            model.estep(viter, vtol, want_elbo=want)               # gpuLDA.jl:356-364, per-document stopping (LDA.jl:171-178)
            model.update_beta()                                    # gpuLDA.jl:365
            model.update_alpha(niter, ntol)                        # gpuLDA.jl:366
            stop = check_elbo(model, checkelbo, printelbo, k, tol)  # gpuLDA.jl:368
        if want and trace is not None:
            trace.append(model.elbo)
        if stop:
            break

    if iter > 0:
        model.update_host()                                    # gpuLDA.jl:373
    model.update_topics()                                      # gpuLDA.jl:374
    return None
