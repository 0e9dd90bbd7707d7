"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

fp64 C restatement (``*_oracle.c``) of the reference's CPU models (src/LDA.jl, src/CTM.jl,
src/CTPF.jl) plus an independent NumPy/SciPy twin (``numpy_twin.py``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this package; nothing under ``topicmodelsvb.jl_b200/`` does.

PARITY UNPINNED: the reference has no tests or golden vectors and Julia is not installed
here, so the oracle is cross-checked only between its two independent restatements.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libtmvb_oracle.so")
_lib = None

_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_c = ctypes


def build(force: bool = False) -> str:
    """Compile the C oracle into oracle/_ref/ (gcc, a few seconds)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = ctypes.CDLL(_SO)
    lib.orc_digamma_export.restype = _c.c_double
    lib.orc_digamma_export.argtypes = [_c.c_double]
    lib.orc_trigamma_export.restype = _c.c_double
    lib.orc_trigamma_export.argtypes = [_c.c_double]
    lib.orc_lgamma_export.restype = _c.c_double
    lib.orc_lgamma_export.argtypes = [_c.c_double]

    lib.orc_lda_train.restype = _c.c_int
    lib.orc_lda_train.argtypes = [
        _c.c_int64, _c.c_int64, _c.c_int64, _i64p, _i64p, _i64p,
        _f64p, _f64p, _f64p, _f64p, _f64p, _f64p,
        _c.c_int, _c.c_double, _c.c_int, _c.c_double, _c.c_int, _c.c_double, _c.c_int,
        _f64p, _i64p, _c.POINTER(_c.c_int), _c.c_int,
    ]
    lib.orc_lda_elbo.restype = _c.c_double
    lib.orc_lda_elbo.argtypes = [
        _c.c_int64, _c.c_int64, _c.c_int64, _i64p, _i64p, _i64p,
        _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _c.c_int,
    ]
    lib.orc_lda_update_alpha.restype = _c.c_int
    lib.orc_lda_update_alpha.argtypes = [_c.c_int64, _c.c_int64, _f64p, _f64p, _c.c_int, _c.c_double]
    lib.orc_lda_estep.restype = _c.c_int64
    lib.orc_lda_estep.argtypes = [
        _c.c_int64, _c.c_int64, _c.c_int64, _i64p, _i64p, _i64p,
        _f64p, _f64p, _f64p, _f64p, _f64p, _c.c_void_p, _c.c_int, _c.c_double, _c.c_int,
    ]
    lib.orc_lda_phi.restype = None
    lib.orc_lda_phi.argtypes = [_c.c_int64, _c.c_int64, _i64p, _i64p, _f64p, _f64p, _f64p]
    lib.orc_ctm_train.restype = _c.c_int
    lib.orc_ctm_train.argtypes = [
        _c.c_int64, _c.c_int64, _c.c_int64, _i64p, _i64p, _i64p,
        _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p,
        _c.c_int, _c.c_double, _c.c_int, _c.c_double, _c.c_int, _c.c_double, _c.c_int,
        _f64p, _i64p, _c.POINTER(_c.c_int), _c.c_int,
    ]
    lib.orc_ctm_elbo.restype = _c.c_double
    lib.orc_ctm_elbo.argtypes = [
        _c.c_int64, _c.c_int64, _c.c_int64, _i64p, _i64p, _i64p,
        _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _c.c_int,
    ]
    lib.orc_ctpf_train.restype = _c.c_int
    lib.orc_ctpf_train.argtypes = ([_c.c_int64] * 4 + [_i64p] * 6 + [_f64p] * 17
                                   + [_c.c_int, _c.c_double, _c.c_int, _c.c_double, _c.c_int, _f64p, _i64p, _c.POINTER(_c.c_int), _c.c_int])
    lib.orc_flda_train.restype = _c.c_int
    lib.orc_flda_train.argtypes = ([_c.c_int64] * 3 + [_i64p] * 3 + [_f64p] * 11
                                   + [_c.c_int, _c.c_double, _c.c_int, _c.c_double, _c.c_int, _c.c_double, _c.c_int, _f64p, _i64p,
                                      _c.POINTER(_c.c_int), _c.c_int])
    lib.orc_flda_elbo.restype = _c.c_double
    lib.orc_flda_elbo.argtypes = [_c.c_int64] * 3 + [_i64p] * 3 + [_c.c_double] + [_f64p] * 10 + [_c.c_int]
    lib.orc_fctm_train.restype = _c.c_int
    lib.orc_fctm_train.argtypes = ([_c.c_int64] * 3 + [_i64p] * 3 + [_c.c_double] + [_f64p] * 13
                                   + [_c.c_int, _c.c_double, _c.c_int, _c.c_double, _c.c_int, _c.c_double, _c.c_int, _f64p, _i64p,
                                      _c.POINTER(_c.c_int), _c.c_int])
    _lib = lib
    return lib


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


class LDAState:
    """The mutable fields of the reference's ``LDA`` struct (LDA.jl:6-23) as fp64 arrays.

    Matrices are stored the way Julia stores them (column-major K x V / K x M), exposed here
    as C-contiguous numpy arrays of shape (V, K) / (M, K): ``beta[j, i]`` is topic i of term j.
    """

    def __init__(self, K, M, V, alpha=None, beta=None, Elogtheta=None, gamma=None):
        from scipy.special import digamma

        self.K, self.M, self.V = int(K), int(M), int(V)
        self.alpha = np.ones(K) if alpha is None else np.array(alpha, dtype=np.float64)
        assert beta is not None, "initial beta is always injected (Julia's RNG stream is not reproducible)"
        self.beta = np.ascontiguousarray(beta, dtype=np.float64).reshape(V, K).copy()
        self.beta_old = self.beta.copy()
        e0 = -np.euler_gamma - digamma(K)  # LDA.jl:38
        self.Elogtheta = (np.full((M, K), e0) if Elogtheta is None
                          else np.ascontiguousarray(Elogtheta, dtype=np.float64).reshape(M, K).copy())
        self.Elogtheta_old = self.Elogtheta.copy()
        self.gamma = (np.ones((M, K)) if gamma is None
                      else np.ascontiguousarray(gamma, dtype=np.float64).reshape(M, K).copy())
        self.elbo = 0.0


def lda_train(st: LDAState, N_cumsum, terms, counts, iter=150, tol=1.0, niter=1000, ntol=None,
              viter=10, vtol=None, checkelbo=1, nthreads=1):
    """train!(model::LDA; ...) (LDA.jl:161-191) on the C oracle.  Returns (elbo_trace, sweeps, iters_done)."""
    lib = load()
    K = st.K
    ntol = 1.0 / K**2 if ntol is None else ntol
    vtol = 1.0 / K**2 if vtol is None else vtol
    N_cumsum = np.ascontiguousarray(N_cumsum, dtype=np.int64)
    terms = np.ascontiguousarray(terms, dtype=np.int64)
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    trace = np.full(iter + 1, np.nan)
    sweeps = np.zeros(max(iter, 1), dtype=np.int64)
    done = _c.c_int(0)
    ce = 0 if (checkelbo is None or checkelbo == float("inf")) else int(checkelbo)
    lib.orc_lda_train(K, st.M, st.V, N_cumsum, terms, counts, st.alpha, st.beta, st.beta_old,
                      st.Elogtheta, st.Elogtheta_old, st.gamma, int(iter), float(tol), int(niter),
                      float(ntol), int(viter), float(vtol), ce, trace, sweeps, _c.byref(done),
                      int(nthreads))
    fin = trace[np.isfinite(trace)]
    if fin.size:
        st.elbo = float(fin[-1])
    return trace, sweeps[:iter], done.value


def lda_elbo(st: LDAState, N_cumsum, terms, counts, nthreads=1) -> float:
    lib = load()
    return lib.orc_lda_elbo(st.K, st.M, st.V, np.ascontiguousarray(N_cumsum, dtype=np.int64),
                            np.ascontiguousarray(terms, dtype=np.int64),
                            np.ascontiguousarray(counts, dtype=np.int64), st.alpha, st.beta,
                            st.beta_old, st.Elogtheta, st.Elogtheta_old, st.gamma, int(nthreads))


def lda_estep(st: LDAState, N_cumsum, terms, counts, viter=10, vtol=None, nthreads=1, want_stats=True):
    """The per-document inner loop of LDA.jl:170-180 for every document.  Returns (stats (V,K) or None, sweeps)."""
    lib = load()
    vtol = 1.0 / st.K**2 if vtol is None else vtol
    stats = np.zeros((st.V, st.K)) if want_stats else None
    sw = lib.orc_lda_estep(st.K, st.M, st.V, np.ascontiguousarray(N_cumsum, dtype=np.int64),
                           np.ascontiguousarray(terms, dtype=np.int64),
                           np.ascontiguousarray(counts, dtype=np.int64), st.alpha, st.beta,
                           st.Elogtheta, st.Elogtheta_old, st.gamma,
                           stats.ctypes.data if stats is not None else None, int(viter), float(vtol),
                           int(nthreads))
    return stats, int(sw)


def lda_update_alpha(K, M, alpha, Elogtheta_sum, niter=1000, ntol=None):
    lib = load()
    a = np.array(alpha, dtype=np.float64)
    ntol = 1.0 / K**2 if ntol is None else ntol
    it = lib.orc_lda_update_alpha(K, M, a, np.ascontiguousarray(Elogtheta_sum, dtype=np.float64), int(niter), float(ntol))
    return a, it


def lda_phi(K, M, N_cumsum, terms, beta_old, Elogtheta_old):
    lib = load()
    N_cumsum = np.ascontiguousarray(N_cumsum, dtype=np.int64)
    phi = np.zeros((int(N_cumsum[-1]), K))
    lib.orc_lda_phi(K, M, N_cumsum, np.ascontiguousarray(terms, dtype=np.int64),
                    np.ascontiguousarray(beta_old, dtype=np.float64),
                    np.ascontiguousarray(Elogtheta_old, dtype=np.float64), phi)
    return phi


class CTMState:
    """The mutable fields of the reference's ``CTM`` struct (CTM.jl:6-25, init :38-49) as fp64 arrays."""

    def __init__(self, K, M, V, beta):
        self.K, self.M, self.V = int(K), int(M), int(V)
        self.mu = np.zeros(K)
        self.sigma = np.eye(K)
        self.invsigma = np.eye(K)
        self.beta = np.ascontiguousarray(beta, dtype=np.float64).reshape(V, K).copy()
        self.beta_old = self.beta.copy()
        self.lam = np.zeros((M, K))
        self.lam_old = np.zeros((M, K))
        self.vsq = np.ones((M, K))
        self.logzeta = np.full(M, 0.5)
        self.elbo = 0.0


def ctm_train(st: CTMState, N_cumsum, terms, counts, iter=150, tol=1.0, niter=1000, ntol=None, viter=10, vtol=None,
              checkelbo=1, nthreads=1):
    """train!(model::CTM; ...) (CTM.jl:185-217) on the C oracle.  Returns (elbo_trace, sweeps, iters_done)."""
    lib = load()
    K = st.K
    ntol = 1.0 / K**2 if ntol is None else ntol
    vtol = 1.0 / K**2 if vtol is None else vtol
    trace = np.full(iter + 1, np.nan)
    sweeps = np.zeros(max(iter, 1), dtype=np.int64)
    done = _c.c_int(0)
    ce = 0 if (checkelbo is None or checkelbo == float("inf")) else int(checkelbo)
    lib.orc_ctm_train(K, st.M, st.V, np.ascontiguousarray(N_cumsum, dtype=np.int64), np.ascontiguousarray(terms, dtype=np.int64),
                      np.ascontiguousarray(counts, dtype=np.int64), st.mu, st.sigma, st.invsigma, st.beta, st.beta_old,
                      st.lam, st.lam_old, st.vsq, st.logzeta, int(iter), float(tol), int(niter), float(ntol), int(viter),
                      float(vtol), ce, trace, sweeps, _c.byref(done), int(nthreads))
    fin = trace[np.isfinite(trace)]
    if fin.size:
        st.elbo = float(fin[-1])
    return trace, sweeps[:iter], done.value


class CTPFState:
    """The mutable fields of the reference's ``CTPF`` struct (CTPF.jl:6-47, init :81-103) as fp64 arrays."""

    def __init__(self, K, M, V, U, alef, hyp=None):
        self.K, self.M, self.V, self.U = int(K), int(M), int(V), int(U)
        self.hyp = np.full(8, 0.1) if hyp is None else np.array(hyp, dtype=np.float64)
        self.alef = np.ascontiguousarray(alef, dtype=np.float64).reshape(V, K).copy()
        self.he = np.ones((max(U, 1), K))[:U] if U else np.ones((0, K))
        self.bet, self.vav, self.dalet, self.het = np.ones(K), np.ones(K), np.ones(K), np.ones(K)
        self.gimel, self.zayin = np.ones((M, K)), np.ones((M, K))
        for n in ("alef", "he", "bet", "vav", "dalet", "het", "gimel", "zayin"):
            setattr(self, n + "_old", getattr(self, n).copy())
        self.elbo = 0.0


def ctpf_train(st: CTPFState, c, iter=150, tol=1.0, viter=10, vtol=None, checkelbo=1, nthreads=1):
    """train!(model::CTPF; ...) (CTPF.jl:344-371) on the C oracle; ``c`` is a synth.CSR with reader lists."""
    lib = load()
    K = st.K
    vtol = 1.0 / K**2 if vtol is None else vtol
    trace = np.full(iter + 1, np.nan)
    sweeps = np.zeros(max(iter, 1), dtype=np.int64)
    done = _c.c_int(0)
    ce = 0 if (checkelbo is None or checkelbo == float("inf")) else int(checkelbo)
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
    he = np.ascontiguousarray(st.he)
    he_old = np.ascontiguousarray(st.he_old)
    lib.orc_ctpf_train(K, st.M, st.V, st.U, i64(c.N_cumsum), i64(c.terms), i64(c.counts), i64(c.R_cumsum), i64(c.readers),
                       i64(c.ratings), st.hyp, st.alef, st.alef_old, he, he_old, st.bet, st.bet_old, st.vav, st.vav_old,
                       st.gimel, st.gimel_old, st.zayin, st.zayin_old, st.dalet, st.dalet_old, st.het, st.het_old,
                       int(iter), float(tol), int(viter), float(vtol), ce, trace, sweeps, _c.byref(done), int(nthreads))
    st.he, st.he_old = he, he_old
    fin = trace[np.isfinite(trace)]
    if fin.size:
        st.elbo = float(fin[-1])
    return trace, sweeps[:iter], done.value


def ctpf_recs(st: CTPFState, c, dtype=np.float64):
    """The recommendation step that ends train!(model::CTPF) (CTPF.jl:381-400; gpuCTPF.jl:709-731 is the same with Float32
    state): scores[d, :] = sum(Eeta .* (Etheta + Eepsilon), dims=1), urecs[u] = findall(ur)[reverse(sortperm(scores[ur, u]))],
    drecs[d] = findall(nr)[reverse(sortperm(scores[d, nr]))].  Plain NumPy; returns (scores (M, U), urecs, drecs) with 1-based
    rankings.  ``dtype`` is the arithmetic type of the scores (float64 = CTPF.jl, float32 = gpuCTPF.jl)."""
    Eeta = (np.asarray(st.he, dtype) / np.asarray(st.vav, dtype)[None, :])                        # (U, K)   CTPF.jl:381
    X = np.asarray(st.gimel, dtype) / np.asarray(st.dalet, dtype)[None, :] + np.asarray(st.zayin, dtype) / np.asarray(st.het, dtype)[None, :]
    scores = np.zeros((st.M, st.U), dtype)
    for d in range(st.M):                                                                         # CTPF.jl:382-386
        scores[d, :] = (Eeta * X[d][None, :]).sum(axis=1)
    Rc, readers = np.asarray(c.R_cumsum), np.asarray(c.readers)
    libs = [[] for _ in range(st.U)]
    for d in range(st.M):
        for u in readers[Rc[d]:Rc[d + 1]]:
            libs[u].append(d)
    urecs, drecs = [], []
    for u in range(st.U):                                                                         # CTPF.jl:388-393
        ur = np.ones(st.M, bool)
        ur[libs[u]] = False
        idx = np.flatnonzero(ur)
        urecs.append(idx[np.argsort(scores[idx, u], kind="stable")[::-1]] + 1)
    for d in range(st.M):                                                                         # CTPF.jl:395-400
        nr = np.ones(st.U, bool)
        nr[readers[Rc[d]:Rc[d + 1]]] = False
        idx = np.flatnonzero(nr)
        drecs.append(idx[np.argsort(scores[d, idx], kind="stable")[::-1]] + 1)
    return scores, urecs, drecs


class FLDAState(LDAState):
    """The mutable fields of the reference's ``fLDA`` struct (fLDA.jl:6-28, init :39-52): LDAState + eta, kappa, tau."""

    def __init__(self, K, M, V, nnz, beta, kappa, alpha=None, eta=0.5):
        super().__init__(K, M, V, alpha=alpha, beta=beta)
        self.eta = np.array([float(eta)])
        self.kappa = np.ascontiguousarray(kappa, dtype=np.float64).copy()
        self.kappa_old = self.kappa.copy()
        self.tau = np.full(int(nnz), float(eta))
        self.tau_old = self.tau.copy()


def flda_train(st: FLDAState, N_cumsum, terms, counts, iter=150, tol=1.0, niter=1000, ntol=None, viter=10, vtol=None,
               checkelbo=1, nthreads=1):
    """train!(model::fLDA; ...) (fLDA.jl:214-247) on the C oracle.  Returns (elbo_trace, sweeps, iters_done)."""
    lib = load()
    K = st.K
    ntol = 1.0 / K**2 if ntol is None else ntol
    vtol = 1.0 / K**2 if vtol is None else vtol
    trace = np.full(iter + 1, np.nan)
    sweeps = np.zeros(max(iter, 1), dtype=np.int64)
    done = _c.c_int(0)
    ce = 0 if (checkelbo is None or checkelbo == float("inf")) else int(checkelbo)
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)  # noqa: E731
    lib.orc_flda_train(K, st.M, st.V, i64(N_cumsum), i64(terms), i64(counts), st.eta, st.alpha, st.kappa, st.kappa_old, st.beta, st.beta_old,
                       st.Elogtheta, st.Elogtheta_old, st.gamma, st.tau, st.tau_old, int(iter), float(tol), int(niter), float(ntol), int(viter),
                       float(vtol), ce, trace, sweeps, _c.byref(done), int(nthreads))
    fin = trace[np.isfinite(trace)]
    if fin.size:
        st.elbo = float(fin[-1])
    return trace, sweeps[:iter], done.value


class FCTMState(CTMState):
    """The mutable fields of the reference's ``fCTM`` struct (fCTM.jl:6-32, init :46-60): CTMState + eta, kappa, tau."""

    def __init__(self, K, M, V, nnz, beta, kappa, eta=0.5):
        super().__init__(K, M, V, beta)
        self.eta = float(eta)
        self.kappa = np.ascontiguousarray(kappa, dtype=np.float64).copy()
        self.kappa_old = self.kappa.copy()
        self.tau = np.full(int(nnz), float(eta))
        self.tau_old = self.tau.copy()


def fctm_train(st: FCTMState, N_cumsum, terms, counts, iter=150, tol=1.0, niter=1000, ntol=None, viter=10, vtol=None, checkelbo=1,
               nthreads=1):
    """train!(model::fCTM; ...) (fCTM.jl:249-290) on the C oracle.  Returns (elbo_trace, sweeps, iters_done)."""
    lib = load()
    K = st.K
    ntol = 1.0 / K**2 if ntol is None else ntol
    vtol = 1.0 / K**2 if vtol is None else vtol
    trace = np.full(iter + 1, np.nan)
    sweeps = np.zeros(max(iter, 1), dtype=np.int64)
    done = _c.c_int(0)
    ce = 0 if (checkelbo is None or checkelbo == float("inf")) else int(checkelbo)
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)  # noqa: E731
    lib.orc_fctm_train(K, st.M, st.V, i64(N_cumsum), i64(terms), i64(counts), float(st.eta), st.mu, st.sigma, st.invsigma, st.kappa,
                       st.kappa_old, st.beta, st.beta_old, st.lam, st.lam_old, st.vsq, st.logzeta, st.tau, st.tau_old, int(iter), float(tol),
                       int(niter), float(ntol), int(viter), float(vtol), ce, trace, sweeps, _c.byref(done), int(nthreads))
    fin = trace[np.isfinite(trace)]
    if fin.size:
        st.elbo = float(fin[-1])
    return trace, sweeps[:iter], done.value
