"""gpufLDA: filtered latent Dirichlet allocation (src/fLDA.jl) on the device.

The reference has no GPU filtered model -- ``@gpu`` leaves fLDA / fCTM untouched (macros.jl:274-278) -- so this mirror follows the
CPU struct (fLDA.jl:6-58) with the conventions of gpuLDA (Float32, K x V / K x M Fortran-ordered matrices); ``tau`` / ``tau_old`` are
flat float32 vectors over the CSR tokens (``tau[d][n]`` of the reference at ``N_cumsum[d] + n``; ``model.tau_of(d)`` slices them).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
from scipy.special import digamma

from . import _lib
from .corpus import Corpus, check_corp
from .dist import Reducer
from .gpu_lda import _fmat


class gpufLDA:
    """GPU accelerated filtered latent Dirichlet allocation model (fields of fLDA.jl:6-28, init :30-58)."""

    def __init__(self, corp: Corpus, K: int, seed: Optional[int] = None, device: int = -1, reducer: Optional[Reducer] = None,
                 M_total: Optional[int] = None, C_total: Optional[float] = None, stream: Optional[int] = None):
        check_corp(corp)
        if not (isinstance(K, (int, np.integer)) and K > 0):
            raise ValueError("number of topics must be a positive integer.")  # fLDA.jl:32
        M, V, _ = corp.size()
        corp = corp.copy()
        flat = corp.flat()
        self.K, self.M, self.V = int(K), int(M), int(V)
        self.N = corp.lengths()             # read-only; rebind (not mutate) to change it
        cs = np.concatenate([[0], np.cumsum(flat.counts)]).astype(np.int64)
        self.C = cs[flat.N_cumsum[1:]] - cs[flat.N_cumsum[:-1]]
        self.corp = corp
        self.topics = [np.arange(1, V + 1) for _ in range(K)]
        rng = np.random.default_rng(seed)
        self.eta = 0.5                                                                   # fLDA.jl:39
        self.alpha = np.ones(K, dtype=np.float32)
        gk = rng.standard_exponential(size=V) if V else np.zeros(0)
        self.kappa = (gk / gk.sum()).astype(np.float32) if V else np.zeros(0, np.float32)  # rand(Dirichlet(V, 1.0)), fLDA.jl:41
        g = rng.standard_exponential(size=(K, V)) if V else np.zeros((K, 0))
        self.beta = np.asfortranarray((g / g.sum(axis=1, keepdims=True)).astype(np.float32)) if V else np.zeros((K, 0), np.float32, order="F")
        e0 = np.float32(-(np.euler_gamma + digamma(K)))                                  # fLDA.jl:47
        self.Elogtheta = np.full((K, M), e0, dtype=np.float32, order="F")
        self.gamma = np.ones((K, M), dtype=np.float32, order="F")
        self.tau = np.full(flat.nnz, self.eta, dtype=np.float32)                         # fLDA.jl:50
        self.kappa_old = self.kappa.copy()
        self.beta_old = self.beta.copy(order="F")
        self.Elogtheta_old = self.Elogtheta.copy(order="F")
        self.tau_old = self.tau.copy()
        self.elbo = 0.0
        self.reducer = reducer
        self.M_total = int(M_total) if M_total is not None else self.M
        self.C_total = float(C_total) if C_total is not None else float(self.C.sum())
        self._device, self._stream = device, stream
        self._h = None
        self._resident = False
        self._pinned = None
        self._corpus_on_device = None

    def tau_of(self, d: int) -> np.ndarray:
        """model.tau[d] of the reference (0-based d)."""
        f = self.corp.flat()
        return self.tau[f.N_cumsum[d]:f.N_cumsum[d + 1]]

    def _handle(self):
        if self._h is None:
            h = C.c_void_p()
            stream = self._stream if self._stream is not None else (self.reducer.stream_ptr() if self.reducer is not None else None)
            _lib.check(_lib.load().tmvb_flda_create(C.byref(h), self.K, self.M, self.V, self._device, stream))
            self._h = h
            from .dist import connect_model_peers
            self._p2p = connect_model_peers(self, "flda")
        return self._h

    def close(self):
        if self._h is not None:
            _lib.load().tmvb_flda_destroy(self._h)
            self._h = None
            self._resident = False
            self._corpus_on_device = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update_buffer(self):
        """What update_buffer!(model) would be for a gpufLDA (cf. modelutils.jl:370-397): flatten, upload corpus and parameters."""
        lib, h = _lib.load(), self._handle()
        f = self.corp.flat()
        z = np.zeros(1, np.int64)
        # an immutable corpus (wrapped from a flattened CSR) is uploaded once per handle, as for gpuLDA
        if not (self.corp.docs is None and self._corpus_on_device is self.corp):
            _lib.check(lib.tmvb_flda_set_corpus(h, _lib.ptr(f.N_cumsum), _lib.ptr(f.terms if f.nnz else z), _lib.ptr(f.counts if f.nnz else z)))
            self._corpus_on_device = self.corp
        self.alpha = np.ascontiguousarray(self.alpha, dtype=np.float32)
        self.kappa = np.ascontiguousarray(self.kappa, dtype=np.float32)
        if self.alpha.shape != (self.K,):
            raise _lib.TopicModelError("alpha must be of length K.")
        if self.kappa.shape != (self.V,):
            raise _lib.TopicModelError("kappa must be of length V")
        self.beta = _fmat(self.beta, self.K, self.V, "beta")
        self.Elogtheta = _fmat(self.Elogtheta, self.K, self.M, "Elogtheta")
        self.gamma = _fmat(self.gamma, self.K, self.M, "gamma")
        self.tau = np.ascontiguousarray(self.tau, dtype=np.float32)
        if self.tau.shape != (f.nnz,):
            raise _lib.TopicModelError("tau must contain one probability per document term.")
        eta = C.c_double(float(self.eta))
        _lib.check(lib.tmvb_flda_upload(h, C.byref(eta), _lib.ptr(self.alpha), _lib.ptr(self.kappa) if self.V else None,
                                        self.beta.ctypes.data if self.V else None, self.Elogtheta.ctypes.data if self.M else None,
                                        self.gamma.ctypes.data if self.M else None, _lib.ptr(self.tau) if f.nnz else None))
        self._resident = True

    def update_host(self):
        """The download half (cf. update_host!, modelutils.jl:501-516): every field of the struct except phi."""
        if not self._resident:
            return
        lib, h = _lib.load(), self._handle()
        K, M, V = self.K, self.M, self.V
        nnz = self.corp.flat().nnz
        eta = C.c_double()
        self.alpha = np.empty(K, np.float32)
        self.kappa, self.kappa_old = np.empty(V, np.float32), np.empty(V, np.float32)
        # page-locked staging arrays, allocated once per model (D2H at PCIe speed): the fields are live views of these buffers and
        # are overwritten by the next update_host! -- copy to keep a snapshot (as for gpuLDA)
        if self._pinned is None:
            pe = _lib.pinned_empty
            self._pinned = dict(beta=pe((K, V), np.float32, order="F"), beta_old=pe((K, V), np.float32, order="F"),
                                Elogtheta=pe((K, M), np.float32, order="F"), Elogtheta_old=pe((K, M), np.float32, order="F"),
                                gamma=pe((K, M), np.float32, order="F"), tau=pe(max(nnz, 1), np.float32)[:nnz],
                                tau_old=pe(max(nnz, 1), np.float32)[:nnz])
        pb = self._pinned
        self.beta, self.beta_old, self.Elogtheta, self.Elogtheta_old = pb["beta"], pb["beta_old"], pb["Elogtheta"], pb["Elogtheta_old"]
        self.gamma, self.tau, self.tau_old = pb["gamma"], pb["tau"], pb["tau_old"]
        hp = lambda a: a.ctypes.data if a.size else None  # noqa: E731
        _lib.check(lib.tmvb_flda_download(h, C.byref(eta), _lib.ptr(self.alpha), hp(self.kappa), hp(self.beta), hp(self.Elogtheta), hp(self.gamma),
                                          hp(self.tau)))
        _lib.check(lib.tmvb_flda_download_old(h, hp(self.kappa_old), hp(self.beta_old), hp(self.Elogtheta_old), hp(self.tau_old)))
        self.eta = float(eta.value)

    def update_topics(self):
        """topics = [reverse(sortperm(vec(beta[i,:]))) for i in 1:K] (fLDA.jl:246)."""
        if not self.V:
            return
        if not self._resident:
            self.topics = [np.argsort(self.beta[i, :], kind="stable")[::-1] + 1 for i in range(self.K)]
            return
        t = np.empty((self.K, self.V), np.int32)
        _lib.check(_lib.load().tmvb_flda_topics(self._handle(), t.ctypes.data))
        self.topics = list(t)

    def stats(self) -> _lib.TmvbStats:
        st = _lib.TmvbStats()
        _lib.check(_lib.load().tmvb_flda_get_stats(self._handle(), C.byref(st)))
        return st

    def estep(self, viter, vtol):
        _lib.check(_lib.load().tmvb_flda_estep(self._handle(), int(viter), float(vtol)))

    def mstep(self, niter, ntol):
        """update_beta!(model), update_kappa!(model), update_alpha!(model, niter, ntol), update_eta!(model) (fLDA.jl:236-239)."""
        lib, h = _lib.load(), self._handle()
        if getattr(self, "_p2p", False):
            _lib.check(lib.tmvb_flda_peer_reduce(h))                            # one kernel over peer memory (tmvb_peer.cu)
        elif self.reducer is not None:
            p = [C.c_void_p() for _ in range(3)]
            n = [C.c_int64() for _ in range(3)]
            _lib.check(lib.tmvb_flda_reduce_buffers(h, C.byref(p[0]), C.byref(n[0]), C.byref(p[1]), C.byref(n[1]), C.byref(p[2]), C.byref(n[2])))
            dev = self.reducer.torch.cuda.current_device()
            self.reducer.allreduce_device([(p[0].value, n[0].value, "<f4"), (p[1].value, n[1].value, "<f4"), (p[2].value, n[2].value, "<f8")], dev)
        _lib.check(lib.tmvb_flda_mstep(h, self.M_total, self.C_total, int(niter), float(ntol)))

    def update_elbo(self) -> float:
        """update_elbo! (fLDA.jl:105-117)."""
        d = C.c_double()
        _lib.check(_lib.load().tmvb_flda_elbo(self._handle(), C.byref(d)))
        v = d.value
        if self.reducer is not None:
            v = self.reducer.allreduce_host(v)
        self.elbo = v
        return self.elbo


def check_model_flda(model: gpufLDA) -> None:
    """check_model(model::fLDA) (modelutils.jl:69-98), minus the element-wise invariants evaluated on the device at upload."""
    E = _lib.TopicModelError
    K, M, V = model.K, model.M, model.V
    f = model.corp.flat()
    if M != len(model.corp):
        raise E("M must equal the number of documents in the corpus.")
    L = model.corp.lengths()
    if model.N is not L and not np.array_equal(model.N, L):
        raise E("N must contain document lengths.")
    if not (0 <= model.eta <= 1):
        raise E("eta must belong to the interval [0,1].")
    a = np.asarray(model.alpha)
    if a.shape != (K,):
        raise E("alpha must be of length K.")
    if not np.all(np.isfinite(a)):
        raise E("alpha must be finite.")
    if not np.all(a > 0):
        raise E("alpha must be positive.")
    if np.shape(model.kappa) != (V,):
        raise E("kappa must be of length V")
    if V and not (np.all(np.asarray(model.kappa) >= 0) and abs(float(np.sum(model.kappa, dtype=np.float64)) - 1.0) < 1e-3):
        raise E("kappa must be a probability vector.")
    if np.shape(model.beta) != (K, V):
        raise E("beta must be of size (K, V).")
    if np.shape(model.Elogtheta) != (K, M):
        raise E("Elogtheta must contain M vectors of length K.")
    if np.shape(model.gamma) != (K, M):
        raise E("gamma must contain M vectors of length K.")
    if not math.isfinite(model.elbo):
        raise E("elbo must be finite.")


def train_flda(model: gpufLDA, iter: int = 150, tol: float = 1.0, niter: int = 1000, ntol: Optional[float] = None, viter: int = 10,
               vtol: Optional[float] = None, checkelbo=1, printelbo: bool = True, trace: Optional[list] = None):
    """train!(model::fLDA; iter, tol, niter, ntol, viter, vtol, checkelbo, printelbo) (fLDA.jl:214-247) with the inner loop on the
    device.  ``trace`` (optional list) receives the ELBO after every checked iteration (slot 0 = the initial update_elbo!)."""
    K = model.K
    ntol = 1.0 / K**2 if ntol is None else ntol
    vtol = 1.0 / K**2 if vtol is None else vtol
    check_model_flda(model)
    if not all(t >= 0 for t in (tol, ntol, vtol)):
        raise ValueError("tolerance parameters must be nonnegative.")
    if not all(t >= 0 for t in (iter, niter, viter)):
        raise ValueError("iteration parameters must be nonnegative.")
    if not ((isinstance(checkelbo, (int, np.integer)) and checkelbo > 0) or checkelbo == math.inf):
        raise ValueError("checkelbo parameter must be a positive integer or Inf.")
    if iter > 0 and viter < 1:
        raise ValueError("viter must be at least 1 (the fused E-step does not keep a stale phi to scatter).")
    if model.corp.flat().nnz == 0 and model.reducer is None:
        iter = 0                                               # fLDA.jl:219
    else:
        model.update_buffer()
    check = checkelbo != math.inf
    if check and checkelbo <= iter:
        model.update_elbo()                                    # fLDA.jl:220
        if trace is not None:
            trace.append(model.elbo)
    for k in range(1, iter + 1):
        model.estep(viter, vtol)                               # fLDA.jl:223-235
        model.mstep(niter, ntol)                               # fLDA.jl:236-239
        if check and k % checkelbo == 0:                       # check_elbo!, modelutils.jl:574-585
            old = model.elbo
            delta = model.update_elbo() - old
            if trace is not None:
                trace.append(model.elbo)
            if printelbo:
                print("%d ∆elbo: %.3f" % (k, delta))
            if delta < tol:
                break
    if iter > 0:
        model.update_host()
    model.update_topics()
    return None
