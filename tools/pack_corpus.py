"""Pack a reference-format docfile (Corpus.jl:277-325: one comma-delimited line per field per
document -- terms, [counts], [readers], [ratings]; 1-based keys) into a compact .npz CSR.

The output (data/_packed/*.npz) is git-ignored: it is derived from the reference's datasets and
travels to the GPU box only through the gpurun snapshot.  Tests and bench.py fall back to
NSF-/CiteULike-shaped synthetic corpora when it is absent.

usage: python tools/pack_corpus.py nsf|citeu [datasets_dir]
"""
import os
import sys

import numpy as np


def pack(docfile, counts=True, readers=False):
    lines_per_doc = 1 + int(counts) + int(readers)
    terms, cnts, rdrs, N, R = [], [], [], [], []
    with open(docfile) as f:
        lines = f.read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    assert len(lines) % lines_per_doc == 0, (len(lines), lines_per_doc)
    for d in range(0, len(lines), lines_per_doc):
        t = np.array(lines[d].split(","), dtype=np.int64) if lines[d] else np.zeros(0, np.int64)
        terms.append(t)
        N.append(len(t))
        k = 1
        if counts:
            cnts.append(np.array(lines[d + k].split(","), dtype=np.int64) if lines[d + k] else np.zeros(0, np.int64))
            k += 1
        if readers:
            r = np.array(lines[d + k].split(","), dtype=np.int64) if lines[d + k] else np.zeros(0, np.int64)
            rdrs.append(r)
            R.append(len(r))
    out = dict(N_cumsum=np.concatenate([[0], np.cumsum(N)]).astype(np.int64),
               terms=(np.concatenate(terms) - 1).astype(np.int32))
    out["counts"] = np.concatenate(cnts).astype(np.int32) if counts else np.ones_like(out["terms"])
    if readers:
        out["R_cumsum"] = np.concatenate([[0], np.cumsum(R)]).astype(np.int64)
        out["readers"] = (np.concatenate(rdrs) - 1).astype(np.int32)
    return out


if __name__ == "__main__":
    which = sys.argv[1]
    root = sys.argv[2] if len(sys.argv) > 2 else "/root/reference/datasets"
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(here, "data", "_packed"), exist_ok=True)
    if which == "nsf":  # readcorp(:nsf): counts=true (Corpus.jl:344)
        o = pack(os.path.join(root, "nsf", "nsfdocs.txt"), counts=True)
        o["V"] = np.int64(sum(1 for _ in open(os.path.join(root, "nsf", "nsfvocab.txt"))))
    elif which == "citeu":  # readcorp(:citeu): counts=true, readers=true (Corpus.jl:351)
        o = pack(os.path.join(root, "citeu", "citeudocs.txt"), counts=True, readers=True)
        o["V"] = np.int64(sum(1 for _ in open(os.path.join(root, "citeu", "citeuvocab.txt"))))
        o["U"] = np.int64(sum(1 for _ in open(os.path.join(root, "citeu", "citeuusers.txt"))))
    else:
        raise SystemExit(__doc__)
    path = os.path.join(here, "data", "_packed", which + ".npz")
    np.savez_compressed(path, **o)
    print(path, {k: (v.shape if hasattr(v, "shape") and v.shape else int(v)) for k, v in o.items()},
          os.path.getsize(path) >> 20, "MiB")
