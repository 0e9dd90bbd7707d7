"""gpuCTM -- host mirror of the reference's ``gpuCTM`` model and its ``train!`` (src/gpuCTM.jl) over the C ABI.

Semantics follow the CPU model (src/CTM.jl): update order phi -> logzeta -> vsq -> lambda inside a sweep,
per-document stopping rule, true log-sum-exp (the OpenCL kernels start their running maximum at 0,
gpuCTM.jl:404,456), lagged-phi ELBO.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np

from . import _lib
from .corpus import Corpus, check_corp
from .dist import Reducer
from .gpu_lda import _fmat


class gpuCTM:
    """GPU accelerated correlated topic model (gpuCTM.jl:6-99).  Matrices are Fortran-ordered float32 (K, V) / (K, M)."""

    def __init__(self, corp: Corpus, K: int, seed: Optional[int] = None, device: int = -1,
                 reducer: Optional[Reducer] = None, M_total: Optional[int] = None, stream: Optional[int] = None):
        check_corp(corp)
        if not (isinstance(K, (int, np.integer)) and K > 0):
            raise ValueError("number of topics must be a positive integer.")  # gpuCTM.jl:55
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
        self.mu = np.zeros(K, dtype=np.float32)                                   # gpuCTM.jl:63
        self.sigma = np.eye(K, dtype=np.float32)                                  # gpuCTM.jl:64
        self.invsigma = np.eye(K, dtype=np.float32)
        g = rng.standard_exponential(size=(K, V)) if V else np.zeros((K, 0))
        self.beta = np.asfortranarray((g / g.sum(axis=1, keepdims=True)).astype(np.float32)) if V else np.zeros((K, 0), np.float32, order="F")
        self.lam = np.zeros((K, M), dtype=np.float32, order="F")                  # `lambda`, gpuCTM.jl:67
        self.vsq = np.ones((K, M), dtype=np.float32, order="F")                   # gpuCTM.jl:69
        self.logzeta = np.full(M, 0.5, dtype=np.float32)                          # gpuCTM.jl:70
        self.beta_old = self.beta.copy(order="F")
        self.lam_old = self.lam.copy(order="F")
        self.elbo = 0.0
        self.reducer = reducer
        self.M_total = int(M_total) if M_total is not None else self.M
        self._device, self._stream = device, stream
        self._h = None
        self._resident = False
        self._pinned = None

    # the reference's field is called `lambda`
    def __getattr__(self, name):
        if name == "lambda_":
            return self.lam
        raise AttributeError(name)

    def _handle(self):
        if self._h is None:
            lib = _lib.load()
            h = C.c_void_p()
            stream = self._stream if self._stream is not None else (self.reducer.stream_ptr() if self.reducer is not None else None)
            _lib.check(lib.tmvb_ctm_create(C.byref(h), self.K, self.M, self.V, self._device, stream))
            self._h = h
            from .dist import connect_model_peers
            self._p2p = connect_model_peers(self, "ctm")
        return self._h

    def close(self):
        if self._h is not None:
            _lib.load().tmvb_ctm_destroy(self._h)
            self._h = None
            self._resident = False
            self._corpus_on_device = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update_buffer(self):
        """update_buffer!(model::gpuCTM) (modelutils.jl:400-435)."""
        lib, h = _lib.load(), self._handle()
        f = self.corp.flat()
        # an immutable corpus (wrapped from a flattened CSR) is uploaded once per handle, as for gpuLDA
        if not (self.corp.docs is None and getattr(self, "_corpus_on_device", None) is self.corp):
            _lib.check(lib.tmvb_ctm_set_corpus(h, _lib.ptr(f.N_cumsum), _lib.ptr(f.terms), _lib.ptr(f.counts)))
            self._corpus_on_device = self.corp
        self.mu = np.ascontiguousarray(self.mu, dtype=np.float32)
        self.sigma = np.ascontiguousarray(self.sigma, dtype=np.float32)
        if self.sigma.shape != (self.K, self.K):
            raise _lib.TopicModelError("sigma must be of size (K, K).")
        self.beta = _fmat(self.beta, self.K, self.V, "beta")
        self.lam = _fmat(self.lam, self.K, self.M, "lambda")
        self.vsq = _fmat(self.vsq, self.K, self.M, "vsq")
        self.logzeta = np.ascontiguousarray(self.logzeta, dtype=np.float32)
        _lib.check(lib.tmvb_ctm_upload(h, _lib.ptr(self.mu), _lib.ptr(self.sigma), self.beta.ctypes.data, self.lam.ctypes.data,
                                       self.vsq.ctypes.data, _lib.ptr(self.logzeta)))
        self._resident = True

    def update_host(self):
        """update_host!(model::gpuCTM) (modelutils.jl:518-537) minus phi."""
        if not self._resident:
            return
        lib, h = _lib.load(), self._handle()
        if self._pinned is None:
            pe = _lib.pinned_empty
            K, M, V = self.K, self.M, self.V
            self._pinned = dict(beta=pe((K, V), np.float32, order="F"), beta_old=pe((K, V), np.float32, order="F"),
                                lam=pe((K, M), np.float32, order="F"), lam_old=pe((K, M), np.float32, order="F"),
                                vsq=pe((K, M), np.float32, order="F"), topics=pe((K, V), np.int32))
        pb = self._pinned
        self.mu = np.empty(self.K, np.float32)
        self.sigma = np.empty((self.K, self.K), np.float32)
        self.invsigma = np.empty((self.K, self.K), np.float32)
        self.logzeta = np.empty(self.M, np.float32)
        self.beta, self.beta_old, self.lam, self.lam_old, self.vsq = pb["beta"], pb["beta_old"], pb["lam"], pb["lam_old"], pb["vsq"]
        _lib.check(lib.tmvb_ctm_download(h, _lib.ptr(self.mu), _lib.ptr(self.sigma), _lib.ptr(self.invsigma), self.beta.ctypes.data,
                                         self.lam.ctypes.data, self.vsq.ctypes.data, _lib.ptr(self.logzeta)))
        _lib.check(lib.tmvb_ctm_download_old(h, self.beta_old.ctypes.data, self.lam_old.ctypes.data))

    def update_topics(self):
        if not self.V:
            return
        if not self._resident:
            self.topics = [np.argsort(self.beta[i, :], kind="stable")[::-1] + 1 for i in range(self.K)]
            return
        if self._pinned is None:
            self.update_host()
        t = self._pinned["topics"]
        _lib.check(_lib.load().tmvb_ctm_topics(self._handle(), t.ctypes.data))
        self.topics = list(t)

    @property
    def phi(self):
        f = self.corp.flat()
        if not self._resident:
            self.update_buffer()
        out = np.zeros((f.nnz, self.K), dtype=np.float32)
        _lib.check(_lib.load().tmvb_ctm_materialize_phi(self._handle(), _lib.ptr(out)))
        return [out[f.N_cumsum[d]:f.N_cumsum[d + 1]].T for d in range(self.M)]

    def stats(self) -> _lib.TmvbStats:
        st = _lib.TmvbStats()
        _lib.check(_lib.load().tmvb_ctm_get_stats(self._handle(), C.byref(st)))
        return st

    def estep(self, niter, ntol, viter, vtol, want_elbo=True):
        """update_phi!/update_logzeta!/update_vsq!/update_lambda! for v in 1:viter (gpuCTM.jl:497-507) + scatter + moments."""
        _lib.check(_lib.load().tmvb_ctm_estep(self._handle(), int(niter), float(ntol), int(viter), float(vtol), int(bool(want_elbo))))

    def mstep(self):
        """update_beta!(), update_sigma!(), update_mu!() (gpuCTM.jl:509-511)."""
        if getattr(self, "_p2p", False):
            _lib.check(_lib.load().tmvb_ctm_peer_reduce(self._handle()))      # one kernel over peer memory (tmvb_peer.cu)
        elif self.reducer is not None:
            lib, h = _lib.load(), self._handle()
            sp, sn, mp, mn = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_int64()
            _lib.check(lib.tmvb_ctm_reduce_buffers(h, C.byref(sp), C.byref(sn), C.byref(mp), C.byref(mn)))
            dev = self.reducer.torch.cuda.current_device()
            self.reducer.allreduce_device([(sp.value, sn.value, "<f4"), (mp.value, mn.value, "<f8")], dev)
        _lib.check(_lib.load().tmvb_ctm_mstep(self._handle(), self.M_total))

    def update_elbo(self, mode: int = 0) -> float:
        docs, glob = C.c_double(), C.c_double()
        _lib.check(_lib.load().tmvb_ctm_elbo(self._handle(), mode, self.M_total, C.byref(docs), C.byref(glob)))
        d = docs.value
        if mode == 1 and self.reducer is not None:
            d = self.reducer.allreduce_host(d)
        self.elbo = d + glob.value
        return self.elbo


def check_model_ctm(model: gpuCTM) -> None:
    """check_model(model::gpuCTM) (modelutils.jl:281-309): shapes and the global parameters on the host; the
    element-wise invariants of beta / lambda / vsq / logzeta run on the device copy during update_buffer!."""
    E = _lib.TopicModelError
    K, M, V = model.K, model.M, model.V
    if M != len(model.corp):
        raise E("M must equal the number of documents in the corpus.")
    if not np.all(np.isfinite(model.mu)):
        raise E("mu must be finite.")
    if np.shape(model.sigma) != (K, K):
        raise E("sigma must be of size (K, K).")
    try:
        np.linalg.cholesky(np.asarray(model.sigma, dtype=np.float64))
    except np.linalg.LinAlgError:
        raise E("sigma must be positive-definite.")
    if np.shape(model.beta) != (K, V):
        raise E("beta must be of size (K, V).")
    # isstochastic(beta, dims=2) (modelutils.jl:293): on the device copy during update_buffer! (shard_check_stochastic)
    if np.shape(model.lam) != (K, M):
        raise E("lambda must contain M vectors of length K.")
    if np.shape(model.vsq) != (K, M):
        raise E("vsq must contain M vectors of length K.")
    if np.shape(model.logzeta) != (M,):
        raise E("logzeta must be of length M.")
    if not math.isfinite(model.elbo):
        raise E("elbo must be finite.")


def train_ctm(model: gpuCTM, iter: int = 150, tol: float = 1.0, niter: int = 1000, ntol: Optional[float] = None,
              viter: int = 10, vtol: Optional[float] = None, checkelbo=1, printelbo: bool = True, trace: Optional[list] = None):
    """train!(model::gpuCTM; iter, tol, niter, ntol, viter, vtol, checkelbo, printelbo) (gpuCTM.jl:487-519)."""
    from .gpu_lda import check_elbo

    K = model.K
    ntol = 1.0 / K**2 if ntol is None else ntol
    vtol = 1.0 / K**2 if vtol is None else vtol
    check_model_ctm(model)
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
    if model.corp.flat().nnz == 0 and model.reducer is None:
        iter = 0
    else:
        model.update_buffer()
    check = checkelbo != math.inf
    if check and checkelbo <= iter:
        model.update_elbo(1)
        if trace is not None:
            trace.append(model.elbo)
    for k in range(1, iter + 1):
        want = check and (k % checkelbo == 0)
        model.estep(niter, ntol, viter, vtol, want_elbo=want)       # gpuCTM.jl:497-507
        model.mstep()                                               # gpuCTM.jl:509-511
        stop = check_elbo(model, checkelbo, printelbo, k, tol)      # gpuCTM.jl:513
        if want and trace is not None:
            trace.append(model.elbo)
        if stop:
            break
    if iter > 0:
        model.update_host()
    model.update_topics()                                           # gpuCTM.jl:517
    return None


class gpufCTM(gpuCTM):
    """GPU accelerated filtered correlated topic model: the fields of fCTM.jl:6-32 (init :34-64) on top of gpuCTM -- ``eta``,
    ``kappa`` (V), ``tau`` / ``tau_old`` (flat float32 over the CSR tokens; ``tau_of(d)`` = the reference's ``tau[d]``).  The reference
    has no GPU version (macros.jl:277-278 skips fCTM); ``train(model, ...)`` follows train!(::fCTM) (fCTM.jl:249-290)."""

    def __init__(self, corp: Corpus, K: int, seed: Optional[int] = None, **kw):
        super().__init__(corp, K, seed=seed, **kw)
        rng = np.random.default_rng(None if seed is None else seed + 1)
        V = self.V
        self.eta = 0.5                                                            # fCTM.jl:46
        gk = rng.standard_exponential(size=V) if V else np.zeros(0)
        self.kappa = (gk / gk.sum()).astype(np.float32) if V else np.zeros(0, np.float32)   # fCTM.jl:50
        self.kappa_old = self.kappa.copy()
        self.tau = np.full(self.corp.flat().nnz, self.eta, dtype=np.float32)      # fCTM.jl:58
        self.tau_old = self.tau.copy()

    def tau_of(self, d: int) -> np.ndarray:
        f = self.corp.flat()
        return self.tau[f.N_cumsum[d]:f.N_cumsum[d + 1]]

    def _handle(self):
        if self._h is None:
            h = C.c_void_p()
            stream = self._stream if self._stream is not None else (self.reducer.stream_ptr() if self.reducer is not None else None)
            _lib.check(_lib.load().tmvb_fctm_create(C.byref(h), self.K, self.M, self.V, self._device, stream))
            self._h = h
            from .dist import connect_model_peers
            self._p2p = connect_model_peers(self, "ctm")       # a filtered handle exports its kappa statistics too
        return self._h

    def update_buffer(self):
        E = _lib.TopicModelError
        if not (0 <= self.eta <= 1):
            raise E("eta must belong to the interval [0,1].")                      # modelutils.jl:104
        self.kappa = np.ascontiguousarray(self.kappa, dtype=np.float32)
        if self.kappa.shape != (self.V,):
            raise E("kappa must be of length V")
        nnz = self.corp.flat().nnz
        self.tau = np.ascontiguousarray(self.tau, dtype=np.float32)
        if self.tau.shape != (nnz,):
            raise E("tau must contain one probability per document term.")
        super().update_buffer()
        eta = C.c_double(float(self.eta))
        _lib.check(_lib.load().tmvb_fctm_upload(self._handle(), C.byref(eta), _lib.ptr(self.kappa) if self.V else None, _lib.ptr(self.tau)))

    def update_host(self):
        if not self._resident:
            return
        super().update_host()
        nnz = self.corp.flat().nnz
        self.kappa, self.kappa_old = np.empty(self.V, np.float32), np.empty(self.V, np.float32)
        if "tau" not in self._pinned:
            self._pinned["tau"] = _lib.pinned_empty(max(nnz, 1), np.float32)[:nnz]
            self._pinned["tau_old"] = _lib.pinned_empty(max(nnz, 1), np.float32)[:nnz]
        self.tau, self.tau_old = self._pinned["tau"], self._pinned["tau_old"]
        hp = lambda a: a.ctypes.data if a.size else None  # noqa: E731
        _lib.check(_lib.load().tmvb_fctm_download(self._handle(), hp(self.kappa), hp(self.kappa_old), hp(self.tau), hp(self.tau_old)))

    @property
    def phi(self):
        raise NotImplementedError("phi of the filtered model is not materialised; rebuild it from tau_old / beta_old / lam_old (fCTM.jl:125)")

    def mstep(self):
        """update_beta!(), update_kappa!(), update_sigma!(), update_mu!() (fCTM.jl:274-277)."""
        self._handle()
        if self.reducer is not None and not getattr(self, "_p2p", False):
            lib, h = _lib.load(), self._handle()
            kp, kn = C.c_void_p(), C.c_int64()
            _lib.check(lib.tmvb_fctm_reduce_buffers(h, C.byref(kp), C.byref(kn)))
            self.reducer.allreduce_device([(kp.value, kn.value, "<f4")], self.reducer.torch.cuda.current_device())
        super().mstep()

    def update_elbo(self, mode: int = 1) -> float:
        """update_elbo! (fCTM.jl:120-130): always the stand-alone evaluation of the device state."""
        return super().update_elbo(1)
