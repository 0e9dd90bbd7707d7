"""gpuCTPF -- host mirror of the reference's ``gpuCTPF`` model and its ``train!`` (src/gpuCTPF.jl) over the C ABI.

Semantics follow the CPU model (src/CTPF.jl): ``vav`` in the xi update (the OpenCL kernel has ``bet``,
gpuCTPF.jl:624 vs CTPF.jl:336), per-document stopping rule, lagged phi/xi ELBO.
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

_HYP = "abcdefgh"


class gpuCTPF:
    """GPU accelerated collaborative topic Poisson factorization model (gpuCTPF.jl:6-153).
    Matrices are Fortran-ordered float32: alef (K, V), he (K, U), gimel / zayin (K, M)."""

    def __init__(self, corp: Corpus, K: int, seed: Optional[int] = None, device: int = -1,
                 reducer: Optional[Reducer] = None, M_total: Optional[int] = None, stream: Optional[int] = None):
        check_corp(corp)
        if not (isinstance(K, (int, np.integer)) and K > 0):
            raise ValueError("number of topics must be a positive integer.")  # gpuCTPF.jl:78
        M, V, U = corp.size()
        corp = corp.copy()      # the reference stores copy(corp): later edits of the caller's corpus do not reach the model
        flat = corp.flat()
        self.K, self.M, self.V, self.U = int(K), int(M), int(V), int(U)
        self.N = corp.lengths()             # read-only; rebind (not mutate) to change it
        cs = np.concatenate([[0], np.cumsum(flat.counts)]).astype(np.int64)
        self.C = cs[flat.N_cumsum[1:]] - cs[flat.N_cumsum[:-1]]
        self.R = np.diff(flat.R_cumsum).astype(np.int64) if flat.R_cumsum is not None else np.zeros(M, np.int64)
        self.corp = corp
        self.topics = [np.arange(1, V + 1) for _ in range(K)]
        self.a = self.b = self.c = self.d = self.e = self.f = self.g = self.h = 0.1     # gpuCTPF.jl:107
        rng = np.random.default_rng(seed)
        gm = rng.standard_exponential(size=(K, V)) if V else np.zeros((K, 0))
        dirichlet = gm / gm.sum(axis=1, keepdims=True) if V else gm
        self.alef = np.asfortranarray(np.exp(dirichlet - 0.5).astype(np.float32))       # gpuCTPF.jl:109
        self.he = np.ones((K, U), dtype=np.float32, order="F")
        self.bet, self.vav, self.dalet, self.het = (np.ones(K, dtype=np.float32) for _ in range(4))
        self.gimel = np.ones((K, M), dtype=np.float32, order="F")
        self.zayin = np.ones((K, M), dtype=np.float32, order="F")
        for n in ("alef", "he", "bet", "vav", "dalet", "het", "gimel", "zayin"):
            setattr(self, n + "_old", getattr(self, n).copy(order="F") if getattr(self, n).ndim == 2 else getattr(self, n).copy())
        self.elbo = 0.0
        self.reducer = reducer
        self.M_total = int(M_total) if M_total is not None else self.M
        self._device, self._stream = device, stream
        self._h = None
        self._resident = False
        self._pinned = None

    def _handle(self):
        if self._h is None:
            lib = _lib.load()
            h = C.c_void_p()
            stream = self._stream if self._stream is not None else (self.reducer.stream_ptr() if self.reducer is not None else None)
            _lib.check(lib.tmvb_ctpf_create(C.byref(h), self.K, self.M, self.V, self.U, self._device, stream))
            self._h = h
            from .dist import connect_model_peers
            self._p2p = connect_model_peers(self, "ctpf")
        return self._h

    def close(self):
        if self._h is not None:
            _lib.load().tmvb_ctpf_destroy(self._h)
            self._h = None
            self._resident = False
            self._corpus_on_device = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _hyp(self):
        return np.array([getattr(self, n) for n in _HYP], dtype=np.float64)

    def update_buffer(self):
        """update_buffer!(model::gpuCTPF) (modelutils.jl:438-494)."""
        lib, h = _lib.load(), self._handle()
        f = self.corp.flat()
        z = np.zeros(1, np.int64)
        rc = f.R_cumsum if f.R_cumsum is not None else np.zeros(self.M + 1, np.int64)
        rd = f.readers if f.readers is not None and len(f.readers) else z
        rt = f.ratings if f.ratings is not None and len(f.ratings) else z
        # an immutable corpus (wrapped from a flattened CSR) is uploaded once per handle, as for gpuLDA
        if not (self.corp.docs is None and getattr(self, "_corpus_on_device", None) is self.corp):
            _lib.check(lib.tmvb_ctpf_set_corpus(h, _lib.ptr(f.N_cumsum), _lib.ptr(f.terms if f.nnz else z), _lib.ptr(f.counts if f.nnz else z),
                                                _lib.ptr(np.ascontiguousarray(rc)), _lib.ptr(np.ascontiguousarray(rd)), _lib.ptr(np.ascontiguousarray(rt))))
            self._corpus_on_device = self.corp
        self.alef = _fmat(self.alef, self.K, self.V, "alef")
        self.he = _fmat(self.he, self.K, self.U, "he")
        self.gimel = _fmat(self.gimel, self.K, self.M, "gimel")
        self.zayin = _fmat(self.zayin, self.K, self.M, "zayin")
        for n in ("bet", "vav", "dalet", "het"):
            v = np.ascontiguousarray(getattr(self, n), dtype=np.float32)
            if v.shape != (self.K,):
                raise _lib.TopicModelError("%s must be of length K." % n)
            setattr(self, n, v)
        hyp = self._hyp()
        _lib.check(lib.tmvb_ctpf_upload(h, _lib.ptr(hyp), self.alef.ctypes.data, self.he.ctypes.data if self.U else None,
                                        _lib.ptr(self.bet), _lib.ptr(self.vav), self.gimel.ctypes.data, self.zayin.ctypes.data,
                                        _lib.ptr(self.dalet), _lib.ptr(self.het)))
        self._resident = True

    def update_host(self):
        """update_host!(model::gpuCTPF) (modelutils.jl:540-570) minus phi / xi."""
        if not self._resident:
            return
        lib, h = _lib.load(), self._handle()
        K, M, V, U = self.K, self.M, self.V, self.U
        if self._pinned is None:
            pe = _lib.pinned_empty
            self._pinned = {n: pe((K, V), np.float32, order="F") for n in ("alef", "alef_old")}
            self._pinned.update({n: pe((K, max(U, 1)), np.float32, order="F")[:, :U] for n in ("he", "he_old")})
            self._pinned.update({n: pe((K, M), np.float32, order="F") for n in ("gimel", "gimel_old", "zayin", "zayin_old")})
            self._pinned["topics"] = pe((K, V), np.int32)
        pb = self._pinned
        for n in ("alef", "he", "gimel", "zayin"):
            setattr(self, n, pb[n])
            setattr(self, n + "_old", pb[n + "_old"])
        for n in ("bet", "vav", "dalet", "het"):
            setattr(self, n, np.empty(K, np.float32))
            setattr(self, n + "_old", np.empty(K, np.float32))
        hp = lambda a: a.ctypes.data if a.size else None
        _lib.check(lib.tmvb_ctpf_download(h, hp(self.alef), hp(self.he), _lib.ptr(self.bet), _lib.ptr(self.vav), hp(self.gimel), hp(self.zayin),
                                          _lib.ptr(self.dalet), _lib.ptr(self.het)))
        _lib.check(lib.tmvb_ctpf_download_old(h, hp(self.alef_old), hp(self.he_old), _lib.ptr(self.bet_old), _lib.ptr(self.vav_old),
                                              hp(self.gimel_old), hp(self.zayin_old), _lib.ptr(self.dalet_old), _lib.ptr(self.het_old)))

    def update_topics(self):
        """topics = ranking of Ebeta = alef ./ bet per topic (gpuCTPF.jl:706-707)."""
        if not self.V:
            return
        if not self._resident:
            self.topics = [np.argsort(self.alef[i, :], kind="stable")[::-1] + 1 for i in range(self.K)]
            return
        if self._pinned is None:
            self.update_host()
        t = self._pinned["topics"]
        _lib.check(_lib.load().tmvb_ctpf_topics(self._handle(), t.ctypes.data))
        self.topics = list(t)

    def scores(self) -> np.ndarray:
        """scores[d, u] = sum_i Eeta[i,u] (Etheta[i,d] + Eepsilon[i,d]) (gpuCTPF.jl:709-714); host-side, on demand."""
        Eeta = self.he / self.vav[:, None]
        Eth = self.gimel / self.dalet[:, None] + self.zayin / self.het[:, None]
        return (Eth.T.astype(np.float32) @ Eeta.astype(np.float32))

    # ---- recommendations (gpuCTPF.jl:88-105, 709-731): the reference materialises scores (M x U) and the two complete
    # rankings (M*U - sum R entries each, ~0.75 GB of Int at CiteULike) at the end of every train!; here one row / column of
    # the score matrix is formed and ranked when it is asked for (showdrecs / showurecs look at a handful of them).
    @property
    def libs(self):
        """libs[u] = 1-based documents user u+1 has in their library (gpuCTPF.jl:88-91)."""
        if getattr(self, "_libs", None) is None:
            f = self.corp.flat()
            if f.readers is None or not len(f.readers):
                self._libs = [np.zeros(0, np.int64) for _ in range(self.U)]
            else:
                doc_of = np.repeat(np.arange(1, self.M + 1), np.diff(f.R_cumsum))
                order = np.argsort(f.readers, kind="stable")
                cut = np.searchsorted(f.readers[order], np.arange(self.U + 1))
                self._libs = [doc_of[order[cut[u]:cut[u + 1]]] for u in range(self.U)]
        return self._libs

    def _Etheta_eps(self):
        return self.gimel / self.dalet[:, None] + self.zayin / self.het[:, None]      # (K, M): Etheta + Eepsilon

    def _rank(self, vals, excluded_1based, n):
        keep = np.ones(n, dtype=bool)
        keep[np.asarray(excluded_1based, dtype=np.int64) - 1] = False
        idx = np.flatnonzero(keep)
        return idx[np.argsort(vals[idx], kind="stable")[::-1]] + 1                    # findall(.)[reverse(sortperm(.))], 1-based

    @property
    def drecs(self):
        """drecs[d] = users who have not read document d+1, by descending score (gpuCTPF.jl:724-729); ranked on access."""
        model = self

        class _D:
            def __len__(self):
                return model.M

            def __getitem__(self, d):
                f = model.corp.flat()
                row = (model._Etheta_eps()[:, d].astype(np.float32) @ (model.he / model.vav[:, None]).astype(np.float32))
                readers = f.readers[f.R_cumsum[d]:f.R_cumsum[d + 1]] + 1 if f.readers is not None else []
                return model._rank(row, readers, model.U)
        return _D()

    @property
    def urecs(self):
        """urecs[u] = documents not in user u+1's library, by descending score (gpuCTPF.jl:716-722); ranked on access."""
        model = self

        class _U:
            def __len__(self):
                return model.U

            def __getitem__(self, u):
                col = model._Etheta_eps().T.astype(np.float32) @ (model.he[:, u] / model.vav).astype(np.float32)
                return model._rank(col, model.libs[u], model.M)
        return _U()

    def update_recs(self, scores: bool = True, urecs: bool = True, drecs: bool = True, mode: int = 0):
        """The recommendation step that ends train!(::gpuCTPF) (gpuCTPF.jl:709-731) on the device (tmvb_ctpf_recs): the dense
        M x U score matrix on the tensor cores and both complete rankings by segmented sorts.  Returns (scores, urecs, drecs):
        scores an (M, U) float32 view (Fortran order: Julia's model.scores), urecs / drecs lists of 1-based int32 arrays
        (views into one concatenated buffer each).  Outputs not asked for are None.  One GPU only: a sharded model has no
        process that holds every document."""
        if self.reducer is not None:
            raise _lib.TopicModelError("update_recs needs the whole corpus on one GPU")
        if not self._resident:
            self.update_buffer()
        M, U = self.M, self.U
        total = M * U - (int(self.corp.flat().R_cumsum[-1]) if M else 0)
        sc = np.empty((M, U), dtype=np.float32, order="F") if scores else None
        ur = np.empty(max(total, 1), np.int32) if urecs else None
        uo = np.zeros(U + 1, np.int64) if urecs else None
        dr = np.empty(max(total, 1), np.int32) if drecs else None
        do = np.zeros(M + 1, np.int64) if drecs else None
        ptr = lambda a: None if a is None else a.ctypes.data  # noqa: E731
        _lib.check(_lib.load().tmvb_ctpf_recs(self._handle(), ptr(sc), ptr(ur), ptr(uo), ptr(dr), ptr(do), int(mode)))
        ul = [ur[uo[u]:uo[u + 1]] for u in range(U)] if urecs else None
        dl = [dr[do[d]:do[d + 1]] for d in range(M)] if drecs else None
        return sc, ul, dl

    def stats(self) -> _lib.TmvbStats:
        st = _lib.TmvbStats()
        _lib.check(_lib.load().tmvb_ctpf_get_stats(self._handle(), C.byref(st)))
        return st

    def estep(self, viter, vtol, want_elbo=True):
        _lib.check(_lib.load().tmvb_ctpf_estep(self._handle(), int(viter), float(vtol), int(bool(want_elbo))))

    def mstep(self):
        """update_he!(), update_alef!(), update_dalet!(), update_het!(), update_bet!(), update_vav!() (gpuCTPF.jl:699-704)."""
        self._handle()
        if getattr(self, "_p2p", False):
            _lib.check(_lib.load().tmvb_ctpf_peer_reduce(self._handle()))     # one kernel over peer memory (tmvb_peer.cu)
        elif self.reducer is not None:
            lib, h = _lib.load(), self._handle()
            p = [C.c_void_p() for _ in range(3)]
            n = [C.c_int64() for _ in range(3)]
            _lib.check(lib.tmvb_ctpf_reduce_buffers(h, C.byref(p[0]), C.byref(n[0]), C.byref(p[1]), C.byref(n[1]), C.byref(p[2]), C.byref(n[2])))
            dev = self.reducer.torch.cuda.current_device()
            self.reducer.allreduce_device([(p[0].value, n[0].value, "<f4"), (p[1].value, n[1].value, "<f4"), (p[2].value, n[2].value, "<f8")], dev)
        _lib.check(_lib.load().tmvb_ctpf_mstep(self._handle(), self.M_total))

    def update_elbo(self, mode: int = 0) -> float:
        docs, glob = C.c_double(), C.c_double()
        _lib.check(_lib.load().tmvb_ctpf_elbo(self._handle(), mode, self.M_total, C.byref(docs), C.byref(glob)))
        d = docs.value
        if mode == 1 and self.reducer is not None:
            d = self.reducer.allreduce_host(d)
        self.elbo = d + glob.value
        return self.elbo


def check_model_ctpf(model: gpuCTPF) -> None:
    """check_model(model::gpuCTPF) (modelutils.jl:311-360): shapes, hyper-parameters and the K-vectors on the host; the
    element-wise invariants of alef / he / gimel / zayin run on the device copy during update_buffer!."""
    E = _lib.TopicModelError
    K, M, V, U = model.K, model.M, model.V, model.U
    if M != len(model.corp):
        raise E("M must be equal to the number of documents in the corpus.")
    for n in _HYP:
        if not getattr(model, n) > 0:
            raise E("%s must be positive." % n)
    if np.shape(model.alef) != (K, V):
        raise E("alef must be of size (K, V).")
    if np.shape(model.he) != (K, U):
        raise E("he must be of size (K, U)")
    for n in ("bet", "vav", "dalet", "het"):
        v = np.asarray(getattr(model, n))
        if v.shape != (K,):
            raise E("%s must be of length K." % n)
        if not np.all(np.isfinite(v)):
            raise E("%s must be finite." % n)
        if not np.all(v > 0):
            raise E("%s must be positive." % n)
    if np.shape(model.gimel) != (K, M):
        raise E("gimel must contain M vectors of length K.")
    if np.shape(model.zayin) != (K, M):
        raise E("zayin must contain M vectors of length K.")
    if not math.isfinite(model.elbo):
        raise E("elbo must be finite")


def train_ctpf(model: gpuCTPF, iter: int = 150, tol: float = 1.0, viter: int = 10, vtol: Optional[float] = None, checkelbo=1,
               printelbo: bool = True, trace: Optional[list] = None, recs: bool = False):
    """train!(model::gpuCTPF; iter, tol, viter, vtol, checkelbo, printelbo) (gpuCTPF.jl:677-733).  With ``recs=True`` the call
    ends like the reference's: the dense score matrix and the complete drecs / urecs rankings (gpuCTPF.jl:709-731) are formed
    on the device (``model.update_recs()``) and left in ``model.scores_``, ``model.urecs_``, ``model.drecs_``; by default
    they are computed on demand, one row at a time (``model.scores()``, ``model.drecs[d]``, ``model.urecs[u]``)."""
    from .gpu_lda import check_elbo

    K = model.K
    vtol = 1.0 / K**2 if vtol is None else vtol
    check_model_ctpf(model)
    if not all(t >= 0 for t in (tol, vtol)):
        raise ValueError("tolerance parameters must be nonnegative.")
    if not all(t >= 0 for t in (iter, viter)):
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
        model.estep(viter, vtol, want_elbo=want)                    # gpuCTPF.jl:687-697
        model.mstep()                                               # gpuCTPF.jl:699-704
        stop = check_elbo(model, checkelbo, printelbo, k, tol)
        if want and trace is not None:
            trace.append(model.elbo)
        if stop:
            break
    if iter > 0:
        model.update_host()
    model.update_topics()
    if recs:
        model.scores_, model.urecs_, model.drecs_ = model.update_recs()
    return None
