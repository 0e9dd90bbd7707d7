"""Synthetic corpora in the flattened CSR form the device path consumes.

* ``gencorp_lda``  -- draws documents from the LDA generative process the way the reference's
  ``gendoc``/``gencorp`` do (modelutils.jl:594-649), duplicates condensed.  cfg0 of SURVEY.md 8(d).
* ``nsf_shaped``   -- an "NSF-shaped" bag-of-words corpus: document lengths, term popularity and
  count distribution calibrated to the statistics of datasets/nsf (M=128804, V=25319,
  sum(N)=10.45e6, N_d median 76 / p99 188, top-1000 terms carry 59 % of the pairs, 78 % of counts
  are 1).  The benchmark input when the packed real corpus is not on the box.
* ``citeu_shaped`` -- the same for datasets/citeu, with reader lists (CTPF).

All return ``CSR`` tuples with 0-based int64 ids (the layout of modelutils.jl:370-381).
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import numpy as np


class CSR(NamedTuple):
    M: int
    V: int
    N_cumsum: np.ndarray          # int64 [M+1]
    terms: np.ndarray             # int64 [sum N], 0-based
    counts: np.ndarray            # int64 [sum N]
    U: int = 0
    R_cumsum: Optional[np.ndarray] = None   # int64 [M+1]
    readers: Optional[np.ndarray] = None    # int64 [sum R], 0-based
    ratings: Optional[np.ndarray] = None    # int64 [sum R]

    @property
    def nnz(self) -> int:
        return int(self.N_cumsum[-1])

    def shard(self, rank: int, world: int) -> "CSR":
        """Documents d with d % world == rank (the doc -> GPU hash of the north star)."""
        return take_docs(self, np.arange(rank, self.M, world))


def take_docs(c: CSR, idx: np.ndarray) -> CSR:
    idx = np.asarray(idx, dtype=np.int64)

    def _gather(off, *arrs):
        lens = off[idx + 1] - off[idx]
        new_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        pos = np.repeat(off[idx] - new_off[:-1], lens) + np.arange(new_off[-1])
        return new_off, [a[pos] for a in arrs]

    off, (t, cn) = _gather(c.N_cumsum, c.terms, c.counts)
    if c.R_cumsum is not None:
        roff, (rd, rt) = _gather(c.R_cumsum, c.readers, c.ratings)
        return CSR(len(idx), c.V, off, t, cn, c.U, roff, rd, rt)
    return CSR(len(idx), c.V, off, t, cn)


def _condense(doc_ids: np.ndarray, term_ids: np.ndarray, M: int, V: int):
    """Collapse repeated (doc, term) draws into unique pairs with counts."""
    key = doc_ids.astype(np.int64) * V + term_ids.astype(np.int64)
    uk, cnt = np.unique(key, return_counts=True)
    d = uk // V
    N = np.bincount(d, minlength=M)
    return np.concatenate([[0], np.cumsum(N)]).astype(np.int64), (uk % V).astype(np.int64), cnt.astype(np.int64)


def gencorp_lda(M=100, V=500, K=5, seed=0, mean_len=60.0, theta_conc=0.5, topic_conc=0.05) -> CSR:
    rng = np.random.default_rng(seed)
    topics = rng.dirichlet(np.full(V, topic_conc), size=K)          # K x V
    C = np.maximum(rng.poisson(mean_len, size=M), 1)
    doc_ids, term_ids = [], []
    for d in range(M):
        theta = rng.dirichlet(np.full(K, theta_conc))
        z = rng.choice(K, size=C[d], p=theta)
        w = np.array([rng.choice(V, p=topics[k]) for k in z])
        doc_ids.append(np.full(C[d], d))
        term_ids.append(w)
    off, t, c = _condense(np.concatenate(doc_ids), np.concatenate(term_ids), M, V)
    return CSR(M, V, off, t, c)


def _popularity(V: int, shift: float, expo: float) -> np.ndarray:
    w = 1.0 / (np.arange(1, V + 1) + shift) ** expo
    return np.cumsum(w / w.sum())


def _bag_of_words(rng, M, V, mu, sigma, nmax, shift, expo, oversample, geom_p, block=1 << 16):
    """N_d ~ clip(round(LogNormal(mu, sigma)), 1, nmax) target unique terms per document; term ids
    drawn from a shifted-Zipf popularity law (duplicates merged); counts 1 + Geometric."""
    cdf = _popularity(V, shift, expo)
    perm = rng.permutation(V)  # popular terms are not the low ids
    offs, ts, cs = [np.zeros(1, np.int64)], [], []
    base = 0
    for s in range(0, M, block):
        m = min(block, M - s)
        Nt = np.clip(np.rint(rng.lognormal(mu, sigma, size=m)), 1, nmax).astype(np.int64)
        draws = np.maximum(np.rint(Nt * oversample), 1).astype(np.int64)
        d = np.repeat(np.arange(m), draws)
        w = perm[np.minimum(np.searchsorted(cdf, rng.random(d.size)), V - 1)]
        off, t, _ = _condense(d, w, m, V)
        c = 1 + rng.geometric(geom_p, size=t.size) - 1
        offs.append(off[1:] + base)
        base += off[-1]
        ts.append(t)
        cs.append(c.astype(np.int64))
    return np.concatenate(offs), np.concatenate(ts), np.concatenate(cs)


def nsf_shaped(M=128804, V=25319, seed=1) -> CSR:
    rng = np.random.default_rng(seed)
    off, t, c = _bag_of_words(rng, M, V, mu=4.285, sigma=0.42, nmax=396, shift=45.0, expo=1.15,
                              oversample=1.08, geom_p=0.72)
    return CSR(M, V, off, t, c)


def citeu_shaped(M=16980, V=8000, U=5551, seed=2) -> CSR:
    rng = np.random.default_rng(seed)
    off, t, c = _bag_of_words(rng, M, V, mu=4.10, sigma=0.45, nmax=1281, shift=12.0, expo=1.0,
                              oversample=1.14, geom_p=0.75)
    # reader lists: R_d ~ 1 + NegBin-ish heavy tail (mean 12.1, max 321), user popularity Zipf
    R = np.clip(np.rint(rng.lognormal(1.75, 1.0, size=M)), 1, 321).astype(np.int64)
    ucdf = _popularity(U, 30.0, 0.8)
    d = np.repeat(np.arange(M), R)
    u = np.minimum(np.searchsorted(ucdf, rng.random(d.size)), U - 1)
    roff, rd, _ = _condense(d, u, M, U)
    return CSR(M, V, off, t, c, U, roff, rd, np.ones_like(rd))


def load_packed(name: str) -> Optional[CSR]:
    """data/_packed/<name>.npz written by tools/pack_corpus.py from the reference's datasets."""
    import os

    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "_packed", name + ".npz")
    if not os.path.exists(p):
        return None
    z = np.load(p)
    off = z["N_cumsum"].astype(np.int64)
    if "readers" in z:
        rd = z["readers"].astype(np.int64)
        return CSR(len(off) - 1, int(z["V"]), off, z["terms"].astype(np.int64), z["counts"].astype(np.int64),
                   int(z["U"]), z["R_cumsum"].astype(np.int64), rd, np.ones_like(rd))
    return CSR(len(off) - 1, int(z["V"]), off, z["terms"].astype(np.int64), z["counts"].astype(np.int64))


def init_beta(K: int, V: int, seed: int = 7) -> np.ndarray:
    """beta ~ Dirichlet(1_V) per topic (LDA.jl:35), returned (V, K) C-order == Julia's K x V column-major."""
    rng = np.random.default_rng(seed)
    g = rng.standard_exponential(size=(K, V))
    g /= g.sum(axis=1, keepdims=True)
    return np.ascontiguousarray(g.T)


def init_alef(K: int, V: int, seed: int = 7) -> np.ndarray:
    """alef = exp.(rand(Dirichlet(V, 1.0), K)' .- 0.5) (CTPF.jl:83), returned (V, K) C-order."""
    return np.exp(init_beta(K, V, seed) - 0.5)


def gencorp_ctpf(M=60, V=300, U=40, K=4, seed=0, mean_len=40.0, mean_readers=4.0) -> CSR:
    """LDA-generated documents plus random reader lists (ratings all 1, as readcorp(:citeu) yields, Corpus.jl:21,351)."""
    c = gencorp_lda(M=M, V=V, K=K, seed=seed, mean_len=mean_len)
    rng = np.random.default_rng(seed + 1000)
    R = rng.poisson(mean_readers, size=M)
    R[0] = 0                                     # a document nobody has in their library
    d = np.repeat(np.arange(M), R)
    u = rng.integers(0, U, size=d.size)
    roff, rd, _ = _condense(d, u, M, U)
    return CSR(M, V, c.N_cumsum, c.terms, c.counts, U, roff, rd, np.ones_like(rd))
