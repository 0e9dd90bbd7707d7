"""Minimal host-side data model: the output contract of the reference's Corpus.jl that the device
path consumes (Document.terms/counts/readers/ratings, 1-based keys; Corpus.jl:14-26,62-78).

Only what ``update_buffer!`` (modelutils.jl:370-494) reads is mirrored here; corpus cleaning,
vocab handling and text I/O stay with the reference package.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .synth import CSR


class DocumentError(ValueError):
    """Corpus.jl:30-34"""


class CorpusError(ValueError):
    """Corpus.jl:85-89"""


def _ivec(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.int64).reshape(-1))


@dataclass
class Document:
    """Document(;terms, counts, readers, ratings, title) -- keys are 1-based as in the reference."""

    terms: np.ndarray
    counts: Optional[np.ndarray] = None
    readers: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    ratings: Optional[np.ndarray] = None
    title: str = ""

    def __post_init__(self):
        self.terms = _ivec(self.terms)
        self.counts = np.ones_like(self.terms) if self.counts is None else _ivec(self.counts)
        self.readers = _ivec(self.readers)
        self.ratings = np.ones_like(self.readers) if self.ratings is None else _ivec(self.ratings)
        check_doc(self)

    def __len__(self):
        return len(self.terms)


def check_doc(doc: Document) -> None:
    """Corpus.jl:41-49"""
    if not np.all(doc.terms > 0):
        raise DocumentError("all terms must be positive integers.")
    if not np.all(doc.counts > 0):
        raise DocumentError("all counts must be positive integers.")
    if len(doc.terms) != len(doc.counts):
        raise DocumentError("terms and counts vectors must have the same length.")
    if not np.all(doc.readers > 0):
        raise DocumentError("all readers must be positive integers.")
    if not np.all(doc.ratings > 0):
        raise DocumentError("all ratings must be positive integers.")
    if len(doc.readers) != len(doc.ratings):
        raise DocumentError("readers and ratings vectors must have the same length.")


class Corpus:
    """Corpus(;docs, vocab, users).  ``vocab``/``users`` may be sequences of names or plain sizes."""

    def __init__(self, docs: Sequence[Document] = (), vocab=None, users=None):
        self.docs: List[Document] = list(docs)
        self._flat: Optional[CSR] = None
        if isinstance(vocab, (int, np.integer)):
            self.V = int(vocab)
        elif vocab is not None:
            self.V = len(vocab)
        else:
            self.V = int(max([d.terms.max() for d in self.docs if len(d.terms)], default=0))
        if isinstance(users, (int, np.integer)):
            self.U = int(users)
        elif users is not None:
            self.U = len(users)
        else:
            self.U = int(max([d.readers.max() for d in self.docs if len(d.readers)], default=0))
        check_corp(self)

    @classmethod
    def from_csr(cls, c: CSR) -> "Corpus":
        """Wrap a flattened corpus without building per-document objects (large corpora)."""
        self = cls.__new__(cls)
        self.docs = None
        self._flat = c
        self.V, self.U = c.V, c.U
        return self

    def copy(self) -> "Corpus":
        """copy(corp) as the model constructors take it (gpuLDA.jl:84): documents and their vectors are duplicated, so editing
        the caller's corpus after construction does not change what the model trains on.  A corpus wrapped from a flattened
        CSR (`from_csr`, `readcorp`) is immutable by construction (NamedTuple of arrays nobody writes): shared, not duplicated."""
        if self.docs is None:
            return self
        new = Corpus.__new__(Corpus)
        new.docs = [Document(d.terms.copy(), d.counts.copy(), d.readers.copy(), d.ratings.copy(), d.title) for d in self.docs]
        new._flat = None
        new.V, new.U = self.V, self.U
        return new

    def __len__(self):
        return self._flat.M if self.docs is None else len(self.docs)

    def lengths(self) -> np.ndarray:
        """[length(doc) for doc in corp] (the N field of every model, e.g. LDA.jl:30) as ONE read-only Int64 array per flattening:
        the models bind it as ``model.N``, so check_model's "N must contain document lengths" is an identity test unless the
        caller rebinds N (0.25 ms of host time per train! call at NSF size otherwise)."""
        f = self.flat()
        cached = getattr(self, "_lengths", None)
        if cached is None or cached[0] is not f:
            n = np.diff(f.N_cumsum).astype(np.int64)
            n.setflags(write=False)
            self._lengths = cached = (f, n)
        return cached[1]

    def size(self):
        return len(self), self.V, self.U

    def flat(self) -> CSR:
        """The flattening of update_buffer! (modelutils.jl:371-380): 0-based Int64 CSR."""
        if self._flat is not None:
            return self._flat
        M = len(self.docs)
        N = np.array([len(d.terms) for d in self.docs], dtype=np.int64)
        R = np.array([len(d.readers) for d in self.docs], dtype=np.int64)
        cat = lambda xs: (np.concatenate(xs) if len(xs) else np.zeros(0, np.int64)).astype(np.int64)
        self._flat = CSR(M, self.V, np.concatenate([[0], np.cumsum(N)]).astype(np.int64),
                         cat([d.terms for d in self.docs]) - 1, cat([d.counts for d in self.docs]),
                         self.U, np.concatenate([[0], np.cumsum(R)]).astype(np.int64),
                         cat([d.readers for d in self.docs]) - 1, cat([d.ratings for d in self.docs]))
        return self._flat


def _flat32(self):
    """(terms, counts) of flat() narrowed to page-locked Int32, built on first use; (None, None) when a count or term id
    does not fit (the Int64 path is used then)."""
    if getattr(self, "_flat32_cache", None) is None:
        f = self.flat()
        if f.nnz and (int(f.terms.max()) >= 2**31 or int(f.counts.max()) >= 2**31):
            self._flat32_cache = (None, None)
        else:
            from . import _lib
            t32, c32 = getattr(self, "_packed32", None) or (f.terms.astype(np.int32), f.counts.astype(np.int32))
            self._flat32_cache = (_lib.pinned_copy(t32), _lib.pinned_copy(c32))
    return self._flat32_cache


Corpus.flat32 = _flat32


def readcorp(docfile: str, vocabfile: Optional[str] = None, userfile: Optional[str] = None, delim: str = ",", counts: bool = False,
             readers: bool = False, ratings: bool = False, nthreads: int = 0) -> Corpus:
    """readcorp(docfile=..., vocabfile=..., userfile=..., delim, counts, readers, ratings) (Corpus.jl:277-325) through the
    native parser (tmvb_read_docfile): the text is parsed straight into the packed CSR that update_buffer! would build, on all
    host threads.  vocabfile / userfile only size the vocabulary / user set here (one `key<TAB>name` line each, as the
    reference's files); without them V and U are the largest keys seen.  Raises CorpusError with the reference's message when a
    document fails to load."""
    import ctypes as C

    from . import _lib

    lib = _lib.load()
    c = _lib.TmvbCsr()
    rc = lib.tmvb_read_docfile(docfile.encode(), delim.encode()[:1], int(counts), int(readers), int(ratings), int(nthreads), C.byref(c))
    if rc == -6:
        raise CorpusError(lib.tmvb_last_error().decode())
    _lib.check(rc)
    try:
        def arr(ptr, n, dtype):
            return np.ctypeslib.as_array(ptr, shape=(max(int(n), 1),))[: int(n)].astype(dtype) if n else np.zeros(0, dtype)

        M = int(c.M)
        off = np.ctypeslib.as_array(c.N_cumsum, shape=(M + 1,)).copy()
        roff = np.ctypeslib.as_array(c.R_cumsum, shape=(M + 1,)).copy()
        t32, c32 = arr(c.terms, c.nnz, np.int32), arr(c.counts, c.nnz, np.int32)
        r32, g32 = arr(c.readers, c.nr, np.int32), arr(c.ratings, c.nr, np.int32)
        vkeys = _read_keys(vocabfile, "vocab") if vocabfile else None
        ukeys = _read_keys(userfile, "user") if userfile else None
        V = _key_count(vkeys, t32, int(c.max_term), "vocab", "term keys not found in corpus vocabulary")
        U = _key_count(ukeys, r32, int(c.max_reader), "user", "user keys not found in corpus users")
    finally:
        lib.tmvb_free_csr(C.byref(c))
    corp = Corpus.from_csr(CSR(M, V, off, t32.astype(np.int64), c32.astype(np.int64), U, roff, r32.astype(np.int64), g32.astype(np.int64)))
    corp._packed32 = (t32, c32)   # what flat32() pins on first use: no second narrowing pass
    return corp


def _read_keys(path: str, what: str) -> np.ndarray:
    """First column of a `key<TAB>name` file as the reference reads it (readdlm + Dict(zip(keys, names)), Corpus.jl:301-315):
    blank lines are skipped, a repeated key keeps one entry, keys must be positive integers (decimal, or with Julia's
    0x / 0o / 0b prefixes as parse(Int, .) accepts them)."""
    with open(path, "r", encoding="utf-8", errors="replace") as f:
        toks = [ln.split("\t", 1)[0].strip() for ln in f.read().splitlines() if ln.strip()]
    try:
        # the usual file: plain decimal keys, converted in one pass
        if not all(t.isascii() and t.isdigit() for t in toks):
            raise ValueError
        keys = np.array(toks, dtype=np.int64) if toks else np.zeros(0, np.int64)
    except (ValueError, OverflowError):
        keys = []
        for tok in toks:
            try:
                keys.append(int(tok, 0) if tok[:2].lower() in ("0x", "0o", "0b") else int(tok))
            except ValueError:
                try:
                    v = float(tok)          # readdlm yields Float64 for a numeric-looking column; Dict{Int,...} converts integral values
                    if v != int(v):
                        raise ValueError
                    keys.append(int(v))
                except ValueError:
                    raise CorpusError("all %s keys must be positive integers." % what)
    k = np.unique(np.asarray(keys, dtype=np.int64))
    if k.size and k[0] <= 0:
        raise CorpusError("all %s keys must be positive integers." % what)
    return k


def _key_count(keys, used0, max_used1, what, missing_msg) -> int:
    """Size of the vocabulary / user set: the number of distinct keys (length of the reference's Dict), after the check_corp
    invariants (Corpus.jl:111-122): used keys are a subset of the key set; the keys form the unit range 1:length."""
    if keys is None:
        return max_used1
    n = int(keys.size)
    if used0.size:
        if n and int(keys[-1]) == n:
            subset = max_used1 <= n      # keys ARE 1:n (sorted, distinct, positive) and the parser only lets positive keys through
        else:
            present = np.zeros(max(max_used1, int(keys[-1]) if n else 0) + 1, dtype=bool)
            present[keys] = True
            subset = bool(present[1:][used0].all())
        if not subset:
            raise CorpusError("documents contain %s (see fixcorp! function)." % missing_msg)
    if n != (int(keys[-1]) if n else 0):
        raise CorpusError("corpus %s keys must form unit range starting at 1 (see fixcorp! function)." % what)
    return n


def check_corp(corp: Corpus) -> None:
    """Corpus.jl:111-122 (the parts that concern the flattened arrays)."""
    if corp.docs is None:
        return
    for d, doc in enumerate(corp.docs):
        try:
            check_doc(doc)
        except DocumentError:
            raise CorpusError("document %d failed check." % (d + 1))
        if len(doc.terms) and doc.terms.max() > corp.V:
            raise CorpusError("documents contain term keys not found in corpus vocabulary (see fixcorp! function).")
        if len(doc.readers) and doc.readers.max() > corp.U:
            raise CorpusError("documents contain user keys not found in corpus users (see fixcorp! function).")
