"""CPU twin of the library's synthetic-row generator (clip-retrieval_b200/csrc/synth.cu).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Inputs for BASELINE.json's configs are synthetic
(SURVEY.md §8d); the generator is integer arithmetic up to one fp64 sqrt/divide per element so
that the GPU rows and these rows agree in every fp16 bit.
"""
import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)
_CSTEP = np.uint64(0xD1342543DE82EF95)
_L1 = np.uint64(0xA0761D6478BD642F)
_L2 = np.uint64(0x5851F42D4C957F2D)


def _mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _row_key(seed, rows):
    with np.errstate(over="ignore"):
        return _mix64(np.uint64(seed) ^ (np.asarray(rows, dtype=np.uint64) * _G))


def _noise(rkeys, d):
    cols = np.arange(d, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _mix64(rkeys[:, None] + cols[None, :] * _CSTEP)
    b = lambda s: ((h >> np.uint64(s)) & np.uint64(0xFF)).astype(np.int64)  # noqa: E731
    return b(0) + b(8) + b(16) + b(24) - 510


def list_of_rows(centroid_seed, rows, nlist):
    with np.errstate(over="ignore"):
        z = np.uint64(centroid_seed) ^ (np.asarray(rows, dtype=np.uint64) * _L1) ^ _L2
    return (_mix64(z) % np.uint64(nlist)).astype(np.int64)


def rows_int(n, d, row0=0, seed=1234, clustered=False, centroid_seed=7, nlist=0, cw=3, nw=1):
    """Integer-valued rows before normalisation ([n, d] int64)."""
    rows = np.arange(row0, row0 + n, dtype=np.uint64)
    v = _noise(_row_key(seed, rows), d)
    if clustered:
        lists = list_of_rows(centroid_seed, rows, nlist)
        c = _noise(_row_key(centroid_seed, lists.astype(np.uint64)), d)
        v = cw * c + nw * v
    return v


def rows_f32(n, d, row0=0, **spec):
    v = rows_int(n, d, row0, **spec)
    ss = (v * v).sum(axis=1)
    norm = np.sqrt(ss.astype(np.float64))
    norm[ss == 0] = 1.0
    x = (v.astype(np.float64) / norm[:, None]).astype(np.float32)
    x[ss == 0] = 0.0
    return x


def rows_f16(n, d, row0=0, **spec):
    return rows_f32(n, d, row0, **spec).astype(np.float16)


def centroids_f32(nlist, d, centroid_seed=7):
    """The generating centroids of a clustered set, L2-normalised (fp32)."""
    return rows_f32(nlist, d, 0, seed=centroid_seed)
