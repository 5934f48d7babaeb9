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
