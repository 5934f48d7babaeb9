"""ctypes front of oracle/knn_ref.c (TEST INFRASTRUCTURE; see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libknn_ref.so")


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def _lib():
    if not os.path.exists(_SO):
        build()
    lib = C.CDLL(_SO)
    lib.knn_flat_ip_f16.restype = C.c_int
    lib.knn_flat_ip_f16.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_int]
    lib.knn_ivf_ip_f16.restype = C.c_int
    lib.knn_ivf_ip_f16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                   C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return lib


def flat_search(X16, Q32, k, id_base=0, nthreads=0):
    """(D, I, threads_used): exhaustive IP search of fp16 rows on the host cores."""
    X16 = np.ascontiguousarray(X16, dtype=np.float16)
    Q32 = np.ascontiguousarray(Q32, dtype=np.float32)
    nq, d = Q32.shape
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    t = _lib().knn_flat_ip_f16(X16.ctypes.data, X16.shape[0], d, Q32.ctypes.data, nq, k, D.ctypes.data, I.ctypes.data,
                               id_base, nthreads)
    return D, I, t


def ivf_layout(X16, assign, nlist):
    """Rows in list order (stable: ascending id inside a list), offsets [nlist+1], ids per slot — the
    inverted-list layout FAISS holds and b200_index_ivf_lists reports."""
    assign = np.asarray(assign, dtype=np.int64)
    order = np.argsort(assign, kind="stable")
    offsets = np.zeros(nlist + 1, dtype=np.int64)
    np.cumsum(np.bincount(assign, minlength=nlist), out=offsets[1:])
    return np.ascontiguousarray(X16[order]), offsets, order.astype(np.int64)


def ivf_search(Xl16, offsets, ids, C16, Q32, k, nprobe, id_base=0, nthreads=0, return_probes=False):
    """(D, I, threads_used[, probes]): IVF-Flat inner-product search on the host cores over a list-ordered store."""
    Xl16 = np.ascontiguousarray(Xl16, dtype=np.float16)
    C16 = np.ascontiguousarray(C16, dtype=np.float16)
    Q32 = np.ascontiguousarray(Q32, dtype=np.float32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    nq, d = Q32.shape
    nlist = C16.shape[0]
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    probes = np.empty((nq, min(nprobe, nlist)), dtype=np.int64) if return_probes else None
    t = _lib().knn_ivf_ip_f16(Xl16.ctypes.data, offsets.ctypes.data, ids.ctypes.data, C16.ctypes.data, nlist, d,
                              Q32.ctypes.data, nq, k, nprobe, D.ctypes.data, I.ctypes.data,
                              probes.ctypes.data if return_probes else None, nthreads)
    I[I >= 0] += id_base
    return (D, I, t, probes) if return_probes else (D, I, t)
