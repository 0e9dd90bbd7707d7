"""CPU tests of the drop-in boundary: the library builds for sm_100a, loads, exports every symbol
include/tmvb.h declares, and refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tmvb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tmvb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(tm):
    lib = tm._lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libtmvb.so does not export %s" % n
    assert set(names) == set(tm._lib.SIGNATURES), set(names) ^ set(tm._lib.SIGNATURES)
    assert lib.tmvb_version() == 100


def test_header_is_plain_c():
    """The ABI must be consumable from C (Julia ccall / cgo style binders): compile the header with gcc."""
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "tmvb.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_library_is_sm100a_only(tm):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", tm._lib.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_native_sass_evidence(tm):
    """TMA bulk copies + mbarrier transactions + vector reductions are what the E-step is built from."""
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", tm._lib.SO_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UBLKCP", "SYNCS.ARRIVE.TRANS64", "REDG.E.ADD.F32x4", "REDUX.SUM", "LDS.128"):
        assert mnemonic in sass, mnemonic


def test_no_device_fails_loudly(tm):
    lib = tm._lib.load()
    n = C.c_int(-1)
    rc = lib.tmvb_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is visible")
    h = C.c_void_p()
    rc = lib.tmvb_lda_create(C.byref(h), 5, 10, 20, -1, None)
    assert rc != 0 and not h.value
    assert lib.tmvb_last_error()
    c = tm.synth.gencorp_lda(M=10, V=50, K=3, seed=0)
    model = tm.gpuLDA(tm.Corpus.from_csr(c), 3)
    with pytest.raises((tm.TopicModelError, ValueError)):
        tm.train(model, iter=1, printelbo=False)


def test_argument_checks_precede_device_work(tm):
    lib = tm._lib.load()
    h = C.c_void_p()
    assert lib.tmvb_lda_create(C.byref(h), 0, 10, 20, -1, None) < 0          # K > 0 (gpuLDA.jl:47)
    assert b"positive" in lib.tmvb_last_error()
    assert lib.tmvb_lda_create(None, 5, 10, 20, -1, None) < 0
    assert lib.tmvb_lda_estep(None, 10, 0.1, 0) < 0
    assert lib.tmvb_lda_destroy(None) == 0                                  # idempotent on NULL


def test_product_path_never_imports_the_oracle():
    import glob

    for f in glob.glob(os.path.join(ROOT, "topicmodelsvb.jl_b200", "**", "*"), recursive=True):
        if os.path.isfile(f) and f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
            txt = open(f, errors="ignore").read()
            assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f
