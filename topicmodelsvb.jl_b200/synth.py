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


# ---- BASELINE.json configs[4]: the HBM-bound scaling sweep -------------------------------------------------------------
CFG4_M, CFG4_V, CFG4_K, CFG4_BLOCK = 1_000_000, 50_000, 200, 12_500


def _draw_without_replacement(rng, Nt, cdf, V):
    """For every document d, Nt[d] distinct term ids drawn successively from the popularity law `cdf` (a repeated id is
    rejected and redrawn -- successive sampling without replacement), vectorised over the block.  Returns CSR offsets + ids
    (sorted by id inside a document)."""
    m = Nt.size
    have = [np.zeros(0, np.int64)]           # accepted keys d * V + term
    need = Nt.copy()
    got = np.zeros(m, np.int64)
    while True:
        todo = np.nonzero(need > 0)[0]
        if todo.size == 0:
            break
        draws = (need[todo] * 3) // 2 + 8
        d = np.repeat(todo, draws)
        w = np.minimum(np.searchsorted(cdf, rng.random(d.size)), V - 1)
        key = d * V + w
        # first occurrence of every new key, in draw order; drop keys accepted in earlier rounds
        uk, first = np.unique(key, return_index=True)
        if len(have) > 1:
            old = np.concatenate(have)
            keep = ~np.isin(uk, old, assume_unique=True)
            uk, first = uk[keep], first[keep]
        order = np.argsort(first, kind="stable")
        uk, first = uk[order], first[order]
        dd = uk // V
        # rank of each new key among its document's new keys (draw order); keep the first need[d]
        o2 = np.argsort(dd, kind="stable")
        dd_s = dd[o2]
        start = np.searchsorted(dd_s, dd_s, side="left")
        rank = np.arange(dd_s.size) - start
        ok = rank < need[dd_s]
        acc = uk[o2][ok]
        have.append(acc)
        cnt = np.bincount(acc // V, minlength=m)
        need = need - cnt
        got += cnt
    keys = np.sort(np.concatenate(have))
    N = np.bincount(keys // V, minlength=m)
    return np.concatenate([[0], np.cumsum(N)]).astype(np.int64), (keys % V).astype(np.int64)


def cfg4_block(b: int, docs: int = CFG4_BLOCK, V: int = CFG4_V) -> CSR:
    """Block `b` (documents [b*12500, (b+1)*12500)) of the synthetic corpus of SURVEY.md 8(d) cfg4: N_d = clip(round(LogNormal(4.30,
    0.45)), 1, 400) distinct terms per document, ids drawn from a Zipf(1.07) law over V without replacement, counts
    1 + Geometric(0.72); one generator per block (`default_rng([1, b])`) so every rank of a multi-GPU run can build its own
    documents on the box."""
    rng = np.random.default_rng([1, int(b)])
    Nt = np.clip(np.rint(rng.lognormal(4.30, 0.45, size=docs)), 1, 400).astype(np.int64)
    cdf = _popularity(V, 0.0, 1.07)
    off, t = _draw_without_replacement(rng, Nt, cdf, V)
    t = rng.permutation(V)[t]                 # popular terms are not the low ids
    c = rng.geometric(0.72, size=t.size).astype(np.int64)   # support {1, 2, ...} = 1 + Geometric on {0, 1, ...}
    return CSR(docs, V, off, t, c)


def _concat_csr(parts, V) -> CSR:
    base = np.cumsum([0] + [int(p.N_cumsum[-1]) for p in parts])
    off = np.concatenate([np.zeros(1, np.int64)] + [p.N_cumsum[1:] + base[i] for i, p in enumerate(parts)]).astype(np.int64)
    return CSR(sum(p.M for p in parts), V, off, np.concatenate([p.terms for p in parts]), np.concatenate([p.counts for p in parts]))


def _cfg4_block_job(args):
    return cfg4_block(*args)


def cfg4_shard(rank: int = 0, world: int = 1, M: int = CFG4_M, V: int = CFG4_V, procs: int = 0) -> CSR:
    """The documents of rank `rank` of `world` of the cfg4 corpus: blocks b = rank (mod world) of CFG4_BLOCK documents each
    (block-cyclic document -> GPU hash, so a rank generates only what it owns), generated by `procs` worker processes
    (0: one per host core, at most 16)."""
    nb = (M + CFG4_BLOCK - 1) // CFG4_BLOCK
    jobs = [(b, min(CFG4_BLOCK, M - b * CFG4_BLOCK), V) for b in range(rank, nb, world)]
    if procs == 0:
        import os

        try:
            procs = len(os.sched_getaffinity(0))
        except AttributeError:  # pragma: no cover
            procs = os.cpu_count() or 1
        procs = max(1, min(16, procs // max(1, min(world, 8)), len(jobs)))
    if procs > 1 and len(jobs) > 1:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(procs) as pool:
            parts = pool.map(_cfg4_block_job, jobs)
    else:
        parts = [cfg4_block(*j) for j in jobs]
    return _concat_csr(parts, V)


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
