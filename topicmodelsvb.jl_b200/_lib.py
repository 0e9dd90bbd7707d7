"""ctypes binding of libtmvb.so (include/tmvb.h) -- the only way the host mirror reaches the device.

There is no fallback: if the shared library is missing, or no sm_100 device is visible, the
first call raises.  ``build()`` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
SO_PATH = os.path.join(_PKG, "libtmvb.so")
COMM_BLOB_BYTES = 512   # TMVB_COMM_BLOB_BYTES
_lib = None

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class TopicModelError(RuntimeError):
    """Mirror of the reference's TopicModelError (modelutils.jl:1-5)."""


class TmvbCsr(C.Structure):
    """tmvb_csr (include/tmvb.h)"""
    _fields_ = [("M", C.c_int64), ("nnz", C.c_int64), ("nr", C.c_int64), ("max_term", C.c_int64), ("max_reader", C.c_int64),
                ("N_cumsum", C.POINTER(C.c_int64)), ("terms", C.POINTER(C.c_int32)), ("counts", C.POINTER(C.c_int32)),
                ("R_cumsum", C.POINTER(C.c_int64)), ("readers", C.POINTER(C.c_int32)), ("ratings", C.POINTER(C.c_int32))]


class TmvbStats(C.Structure):
    _fields_ = [("estep_ms", C.c_double), ("mstep_ms", C.c_double), ("sweeps", C.c_int64),
                ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


def sources():
    return sorted(glob.glob(os.path.join(_PKG, "csrc", "*.cu")))


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> topicmodelsvb.jl_b200/libtmvb.so"""
    srcs = sources()
    deps = srcs + glob.glob(os.path.join(_PKG, "csrc", "*.cuh")) + [os.path.join(_ROOT, "include", "tmvb.h")]
    if not force and os.path.exists(SO_PATH) and all(os.path.getmtime(d) <= os.path.getmtime(SO_PATH) for d in deps):
        return SO_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for s in srcs:  # compile the translation units in parallel, then link
        o = os.path.splitext(s)[0] + ".o"
        objs.append(o)
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
        if verbose and out:
            print(out)
    cmd = [nvcc] + NVCC_FLAGS + objs + ["-o", SO_PATH]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return SO_PATH


_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/tmvb.h declares
SIGNATURES = {
    "tmvb_version": (C.c_int, []),
    "tmvb_last_error": (C.c_char_p, []),
    "tmvb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "tmvb_alloc_pinned": (C.c_int, [C.POINTER(_vp), C.c_int64]),
    "tmvb_free_pinned": (C.c_int, [_vp]),
    "tmvb_read_docfile": (C.c_int, [C.c_char_p, C.c_char, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(TmvbCsr)]),
    "tmvb_free_csr": (C.c_int, [C.POINTER(TmvbCsr)]),
    "tmvb_lda_create": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, C.c_int64, C.c_int, _vp]),
    "tmvb_lda_destroy": (C.c_int, [_vp]),
    "tmvb_lda_set_corpus": (C.c_int, [_vp, _vp, _vp, _vp]),
    "tmvb_lda_set_corpus32": (C.c_int, [_vp, _vp, _vp, _vp]),
    "tmvb_lda_upload": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "tmvb_lda_set_alpha": (C.c_int, [_vp, _vp]),
    "tmvb_lda_estep": (C.c_int, [_vp, C.c_int, C.c_float, C.c_int]),
    "tmvb_lda_predict": (C.c_int, [_vp, C.c_int, C.c_float]),
    "tmvb_ctm_predict": (C.c_int, [_vp, C.c_int, C.c_float, C.c_int, C.c_float]),
    "tmvb_lda_reduce_buffers": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_int64), C.POINTER(_vp), C.POINTER(C.c_int64)]),
    "tmvb_lda_mstep": (C.c_int, [_vp]),
    "tmvb_lda_comm_export": (C.c_int, [_vp, _vp, C.c_int64]),
    "tmvb_lda_comm_connect": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int64]),
    "tmvb_lda_exchange_mstep": (C.c_int, [_vp]),
    "tmvb_lda_comm_status": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "tmvb_lda_get_elogtheta_sum": (C.c_int, [_vp, _vp]),
    "tmvb_lda_update_alpha": (C.c_int, [_vp, C.c_int64, C.c_int, C.c_double, _vp]),
    "tmvb_lda_iterate": (C.c_int, [_vp, C.c_int, C.c_float, C.c_int, C.c_int64, C.c_int, C.c_double, C.POINTER(C.c_double)]),
    "tmvb_lda_elbo": (C.c_int, [_vp, C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "tmvb_lda_download": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "tmvb_lda_arm_host_mirror": (C.c_int, [_vp, _vp, _vp]),
    "tmvb_lda_download_old": (C.c_int, [_vp, _vp, _vp]),
    "tmvb_lda_materialize_phi": (C.c_int, [_vp, _vp]),
    "tmvb_lda_topics": (C.c_int, [_vp, _vp]),
    "tmvb_lda_sync": (C.c_int, [_vp]),
    "tmvb_lda_get_stats": (C.c_int, [_vp, C.POINTER(TmvbStats)]),
    "tmvb_lda_kld": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "tmvb_ctm_create": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, C.c_int64, C.c_int, _vp]),
    "tmvb_ctm_destroy": (C.c_int, [_vp]),
    "tmvb_ctm_set_corpus": (C.c_int, [_vp, _vp, _vp, _vp]),
    "tmvb_ctm_upload": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tmvb_ctm_estep": (C.c_int, [_vp, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int]),
    "tmvb_ctm_reduce_buffers": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_int64), C.POINTER(_vp), C.POINTER(C.c_int64)]),
    "tmvb_ctm_mstep": (C.c_int, [_vp, C.c_int64]),
    "tmvb_ctm_elbo": (C.c_int, [_vp, C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "tmvb_ctm_download": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tmvb_ctm_download_old": (C.c_int, [_vp, _vp, _vp]),
    "tmvb_ctm_materialize_phi": (C.c_int, [_vp, _vp]),
    "tmvb_ctm_topics": (C.c_int, [_vp, _vp]),
    "tmvb_ctm_get_stats": (C.c_int, [_vp, C.POINTER(TmvbStats)]),
    "tmvb_ctpf_create": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, _vp]),
    "tmvb_ctpf_destroy": (C.c_int, [_vp]),
    "tmvb_ctpf_set_corpus": (C.c_int, [_vp] * 7),
    "tmvb_ctpf_upload": (C.c_int, [_vp] * 10),
    "tmvb_ctpf_estep": (C.c_int, [_vp, C.c_int, C.c_float, C.c_int]),
    "tmvb_ctpf_reduce_buffers": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_int64), C.POINTER(_vp), C.POINTER(C.c_int64), C.POINTER(_vp), C.POINTER(C.c_int64)]),
    "tmvb_ctpf_mstep": (C.c_int, [_vp, C.c_int64]),
    "tmvb_ctpf_elbo": (C.c_int, [_vp, C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "tmvb_ctpf_download": (C.c_int, [_vp] * 9),
    "tmvb_ctpf_download_old": (C.c_int, [_vp] * 9),
    "tmvb_ctpf_topics": (C.c_int, [_vp, _vp]),
    "tmvb_ctpf_recs": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "tmvb_ctpf_get_stats": (C.c_int, [_vp, C.POINTER(TmvbStats)]),
    "tmvb_flda_create": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, C.c_int64, C.c_int, _vp]),
    "tmvb_flda_destroy": (C.c_int, [_vp]),
    "tmvb_flda_set_corpus": (C.c_int, [_vp] * 4),
    "tmvb_flda_upload": (C.c_int, [_vp, C.POINTER(C.c_double)] + [_vp] * 6),
    "tmvb_flda_estep": (C.c_int, [_vp, C.c_int, C.c_float]),
    "tmvb_flda_predict": (C.c_int, [_vp, C.c_int, C.c_float]),
    "tmvb_flda_reduce_buffers": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_int64), C.POINTER(_vp), C.POINTER(C.c_int64), C.POINTER(_vp), C.POINTER(C.c_int64)]),
    "tmvb_flda_mstep": (C.c_int, [_vp, C.c_int64, C.c_double, C.c_int, C.c_double]),
    "tmvb_flda_elbo": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "tmvb_flda_download": (C.c_int, [_vp, C.POINTER(C.c_double)] + [_vp] * 6),
    "tmvb_flda_download_old": (C.c_int, [_vp] * 5),
    "tmvb_flda_topics": (C.c_int, [_vp, _vp]),
    "tmvb_flda_get_stats": (C.c_int, [_vp, C.POINTER(TmvbStats)]),
    "tmvb_ctm_comm_export": (C.c_int, [_vp, _vp, C.c_int64]),
    "tmvb_ctm_comm_connect": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int64]),
    "tmvb_ctm_peer_reduce": (C.c_int, [_vp]),
    "tmvb_ctpf_comm_export": (C.c_int, [_vp, _vp, C.c_int64]),
    "tmvb_ctpf_comm_connect": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int64]),
    "tmvb_ctpf_peer_reduce": (C.c_int, [_vp]),
    "tmvb_flda_comm_export": (C.c_int, [_vp, _vp, C.c_int64]),
    "tmvb_flda_comm_connect": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int64]),
    "tmvb_flda_peer_reduce": (C.c_int, [_vp]),
    "tmvb_fctm_create": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, C.c_int64, C.c_int, _vp]),
    "tmvb_fctm_upload": (C.c_int, [_vp, C.POINTER(C.c_double), _vp, _vp]),
    "tmvb_fctm_download": (C.c_int, [_vp] * 5),
    "tmvb_fctm_reduce_buffers": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_int64)]),
}


def load():
    """dlopen libtmvb.so (no build, no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    so = os.environ.get("TMVB_SO") or SO_PATH   # TMVB_SO: developer A/B runs against another build of the same ABI
    if not os.path.exists(so):
        raise TopicModelError("libtmvb.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                              "there is no CPU fallback")
    lib = C.CDLL(so)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    """0 ok; <0 invalid argument -> ValueError (the reference's ArgumentError); >0 CUDA error -> TopicModelError."""
    if rc == 0:
        return
    msg = load().tmvb_last_error().decode()
    if rc == -5:   # a check_model invariant evaluated on the device
        raise TopicModelError(msg)
    if rc < 0:
        raise ValueError(msg)
    raise TopicModelError("CUDA error %d: %s" % (rc, msg))


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def pinned_empty(shape, dtype, order="C"):
    """numpy array over page-locked host memory (tmvb_alloc_pinned); freed when the array is collected."""
    import weakref

    lib = load()
    dtype = np.dtype(dtype)
    shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(int(x) for x in shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    p = C.c_void_p()
    check(lib.tmvb_alloc_pinned(C.byref(p), nbytes))
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    weakref.finalize(buf, lib.tmvb_free_pinned, p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape, order=order)


def pinned_copy(a, order=None):
    a = np.asarray(a)
    if order is None:
        order = "F" if (a.ndim > 1 and a.flags["F_CONTIGUOUS"] and not a.flags["C_CONTIGUOUS"]) else "C"
    out = pinned_empty(a.shape, a.dtype, order=order)
    out[...] = a
    return out
