"""Writes the inputs of the three small golden cases (tests/golden/{lda_cfg0,ctm_cfg,ctpf_cfg}.npz) in formats the reference
itself reads -- a `readcorp` docfile (Corpus.jl:277-296: terms / counts / readers lines, comma-delimited, 1-based keys) and the
injected initial table as raw little-endian Float64 in Julia's column-major K x V order -- under tests/golden/reference/inputs/.
tools/reference_golden.jl (run by somebody who has Julia + TopicModelsVB.jl) turns them into tests/golden/reference/<case>.json;
tests/test_reference_golden_cpu.py pins the oracle against those dumps when they exist.

usage: python tools/export_reference_inputs.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "reference", "inputs")


def write_docs(path, z, readers):
    off = z["N_cumsum"]
    with open(path, "w") as f:
        for d in range(len(off) - 1):
            a, b = int(off[d]), int(off[d + 1])
            f.write(",".join(str(int(t) + 1) for t in z["terms"][a:b]) + "\n")
            f.write(",".join(str(int(c)) for c in z["counts"][a:b]) + "\n")
            if readers:
                ra, rb = int(z["R_cumsum"][d]), int(z["R_cumsum"][d + 1])
                f.write(",".join(str(int(u) + 1) for u in z["readers"][ra:rb]) + "\n")


def main():
    os.makedirs(OUT, exist_ok=True)
    for case, model, key, iters in (("lda_cfg0", "LDA", "beta0", 20), ("ctm_cfg", "CTM", "beta0", 8), ("ctpf_cfg", "CTPF", "alef0", 8),
                                    ("flda_cfg", "fLDA", "beta0", 8), ("fctm_cfg", "fCTM", "beta0", 6)):
        z = np.load(os.path.join(ROOT, "tests", "golden", case + ".npz"))
        readers = "R_cumsum" in z
        if readers and any(z["R_cumsum"][d + 1] == z["R_cumsum"][d] for d in range(len(z["R_cumsum"]) - 1)):
            # a document nobody reads would be an empty line, which parse(Int, "") rejects: the Julia script re-creates those
            # documents from the explicit length list instead
            pass
        write_docs(os.path.join(OUT, case + "_docs.txt"), z, False)
        K, V = int(z["K"]), int(z["V"])
        table = np.asarray(z[key], dtype=np.float64)            # (V, K) C-order == K x V column-major
        assert table.shape == (V, K)
        table.astype("<f8").tofile(os.path.join(OUT, case + "_init.f64"))
        meta = dict(model=model, K=K, V=V, M=int(len(z["N_cumsum"]) - 1), iter=iters, viter=10, init_field="alef" if model == "CTPF" else "beta")
        if "kappa0" in z:   # the filtered models: the injected initial kappa (fLDA.jl:41 / fCTM.jl:50 draw it from Julia's RNG)
            np.asarray(z["kappa0"], dtype="<f8").tofile(os.path.join(OUT, case + "_kappa.f64"))
        if readers:
            meta["U"] = int(z["U"])
            meta["R_cumsum"] = [int(x) for x in z["R_cumsum"]]
            meta["readers"] = [int(u) + 1 for u in z["readers"]]
        json.dump(meta, open(os.path.join(OUT, case + "_meta.json"), "w"))
        print(case, meta["model"], "K", K, "V", V, "M", meta["M"])


if __name__ == "__main__":
    sys.exit(main())
