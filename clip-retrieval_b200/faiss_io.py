"""Reader (and a minimal writer) of FAISS index files for `load_index` (clip_retrieval/clip_back.py:589-596 reads
`image.index` / `text.index` with `faiss.read_index`).

FAISS is not installable offline and /root/reference holds no index file, so this module restates the published
serialisation of faiss/impl/index_write.cpp / index_read.cpp (faiss-cpu >= 1.7.2, requirements.txt:8) from its
documented layout; it is UNVERIFIED against files written by FAISS itself — the tests round-trip this module's
own writer.  Supported (inner-product metric only — the reference builds `metric_type="ip"` indices):
    IxFI / IxF2 / IxFl   IndexFlat                      -> rows fp32
    IxSQ                 IndexScalarQuantizer, QT_fp16  -> rows fp16
    IwFl                 IndexIVFFlat (quantizer IndexFlat, ArrayInvertedLists)
    IwSq                 IndexIVFScalarQuantizer, QT_fp16, by_residual = false
    IxMp / IxM2          IndexIDMap / IDMap2 around one of the above
Anything else (PQ / OPQ / HNSW codes — what autofaiss picks for billion-scale sets — on-disk inverted lists) raises
NotImplementedError naming the fourcc: those store quantised CODES, which an exact fp16 engine cannot serve
bit-compatibly; re-index from the embedding shards instead (`load_index(<folder of .npy>)`).

Layout (little endian): fourcc u32; header = d i32, ntotal i64, dummy i64 x2, is_trained u8, metric_type i32
(0 = inner product, 1 = L2; metric_arg f32 follows when metric_type > 1); vectors are `size u64` + payload.
"""
import struct

import numpy as np

METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1
QT_FP16 = 4   # faiss::ScalarQuantizer::QuantizerType: 8bit, 4bit, 8bit_uniform, 4bit_uniform, fp16, ...


def _fourcc(s):
    return struct.unpack("<I", s.encode("ascii"))[0]


def _fourcc_str(v):
    return struct.pack("<I", v).decode("ascii", errors="replace")


class _R:
    def __init__(self, buf):
        self.b, self.o = memoryview(buf), 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def vec(self, dtype, unit_bytes=None):
        n = self.take("Q")
        dt = np.dtype(dtype)
        nbytes = n * (unit_bytes if unit_bytes is not None else dt.itemsize)
        a = np.frombuffer(self.b, dtype=np.uint8, count=nbytes, offset=self.o)
        self.o += nbytes
        return a.view(dt)

    def raw(self, dtype, count):
        dt = np.dtype(dtype)
        a = np.frombuffer(self.b, dtype=dt, count=count, offset=self.o)
        self.o += count * dt.itemsize
        return a


def _header(r):
    d, ntotal = r.take("i"), r.take("q")
    r.take("q"); r.take("q")
    trained = r.take("B")
    metric = r.take("i")
    if metric > 1:
        r.take("f")
    return d, ntotal, bool(trained), metric


def _read_sq(r):
    qtype, _rangestat, _arg, d, code_size = r.take("i"), r.take("i"), r.take("f"), r.take("Q"), r.take("Q")
    r.vec(np.float32)   # trained
    if qtype != QT_FP16:
        raise NotImplementedError("ScalarQuantizer type %d: only QT_fp16 (4) stores the rows an fp16 engine serves" % qtype)
    if code_size != 2 * d:
        raise ValueError("QT_fp16 code size %d for d=%d" % (code_size, d))
    return d


def _read_invlists(r, nlist_expected, code_dtype, d):
    tag = _fourcc_str(r.take("I"))
    if tag == "il00":
        raise ValueError("index has no inverted lists")
    if tag != "ilar":
        raise NotImplementedError("inverted lists %r: only in-memory ArrayInvertedLists ('ilar') are readable "
                                  "(on-disk lists keep their codes in a separate .ivfdata file)" % tag)
    nlist, code_size = r.take("Q"), r.take("Q")
    if nlist != nlist_expected or code_size != d * np.dtype(code_dtype).itemsize:
        raise ValueError("inverted lists: nlist %d (expected %d), code size %d for d=%d" % (nlist, nlist_expected, code_size, d))
    fmt = _fourcc_str(r.take("I"))
    sizes = np.zeros(nlist, dtype=np.int64)
    if fmt == "full":
        sizes[:] = r.vec(np.uint64).astype(np.int64)
    elif fmt == "sprs":
        pairs = r.vec(np.uint64).astype(np.int64).reshape(-1, 2)
        sizes[pairs[:, 0]] = pairs[:, 1]
    else:
        raise ValueError("inverted lists: unknown size encoding %r" % fmt)
    rows, ids = [], []
    for n in sizes:
        if n > 0:
            rows.append(r.raw(code_dtype, int(n) * d).reshape(int(n), d))
            ids.append(r.raw(np.int64, int(n)))
    rows = np.concatenate(rows) if rows else np.zeros((0, d), code_dtype)
    ids = np.concatenate(ids) if ids else np.zeros(0, np.int64)
    return sizes, rows, ids


def _read_index(r):
    tag = _fourcc_str(r.take("I"))
    if tag in ("IxFI", "IxF2", "IxFl"):
        d, ntotal, _, metric = _header(r)
        xb = r.vec(np.float32)           # codes as bytes / 4 == floats: the count is in 4-byte units either way
        return {"kind": "flat", "d": d, "ntotal": ntotal, "metric": metric, "rows": xb.reshape(ntotal, d), "ids": None}
    if tag == "IxSQ":
        d, ntotal, _, metric = _header(r)
        _read_sq(r)
        codes = r.vec(np.uint8)
        return {"kind": "flat", "d": d, "ntotal": ntotal, "metric": metric, "rows": codes.view(np.float16).reshape(ntotal, d), "ids": None}
    if tag in ("IwFl", "IwSq"):
        d, ntotal, _, metric = _header(r)
        nlist, nprobe = r.take("Q"), r.take("Q")
        quant = _read_index(r)
        if quant["kind"] != "flat":
            raise NotImplementedError("IVF coarse quantiser %r: only a flat quantiser is supported" % quant["kind"])
        dm_type = r.take("b")
        r.vec(np.int64)                  # direct map array
        if dm_type == 2:
            raise NotImplementedError("IVF direct map stored as a hash table")
        if tag == "IwSq":
            _read_sq(r)
            r.take("Q")                  # code_size
            if r.take("B"):
                raise NotImplementedError("IVF scalar quantiser with by_residual=true stores residual codes")
            dt = np.float16
        else:
            r.take("Q")                  # code_size
            dt = np.float32
        sizes, rows, ids = _read_invlists(r, nlist, dt, d)
        return {"kind": "ivfflat", "d": d, "ntotal": ntotal, "metric": metric, "nlist": int(nlist), "nprobe": int(nprobe),
                "centroids": np.asarray(quant["rows"], dtype=np.float32), "list_sizes": sizes, "rows": rows, "ids": ids}
    if tag in ("IxMp", "IxM2"):
        d, ntotal, _, metric = _header(r)
        sub = _read_index(r)
        id_map = r.vec(np.int64)
        if sub["kind"] == "flat":
            sub["ids"] = np.array(id_map)
        else:
            sub["ids"] = np.asarray(id_map)[sub["ids"]]
        return sub
    raise NotImplementedError(
        "FAISS index type %r is not readable: it stores quantised codes or a graph (PQ / OPQ / HNSW / pre-transform "
        "chains, which autofaiss selects for large sets); this engine serves exact fp16 rows — rebuild from the "
        "embedding shards with load_index(<folder of .npy>) or build_ivf_index" % tag)


def read_faiss_index(path):
    """Parse a FAISS index file into plain arrays (see the module docstring for the supported types)."""
    with open(path, "rb") as f:
        buf = f.read()
    info = _read_index(_R(buf))
    if info["metric"] != METRIC_INNER_PRODUCT:
        raise NotImplementedError("FAISS index with metric %d: the search path is inner product (cosine on normalised rows)" % info["metric"])
    return info


def looks_like_faiss_index(path):
    try:
        with open(path, "rb") as f:
            head = f.read(4)
    except OSError:
        return False
    return len(head) == 4 and head[:2] in (b"Ix", b"Iw", b"IH", b"Ib", b"IB", b"IR", b"I2", b"IL")


# ---- writer (fixtures for the tests; the same layout) -----------------------------------------------------
def _w_header(out, d, ntotal, metric):
    out.append(struct.pack("<iqqqBi", d, ntotal, 1 << 20, 1 << 20, 1, metric))


def _w_vec(out, a):
    a = np.ascontiguousarray(a)
    out.append(struct.pack("<Q", a.size))
    out.append(a.tobytes())


def write_flat(path, rows, metric=METRIC_INNER_PRODUCT, fp16=False, id_map=None):
    rows = np.ascontiguousarray(rows, dtype=np.float16 if fp16 else np.float32)
    n, d = rows.shape
    out = []
    if id_map is not None:
        out.append(struct.pack("<I", _fourcc("IxMp")))
        _w_header(out, d, n, metric)
    if fp16:
        out.append(struct.pack("<I", _fourcc("IxSQ")))
        _w_header(out, d, n, metric)
        out.append(struct.pack("<iifQQ", QT_FP16, 0, 0.0, d, 2 * d))
        _w_vec(out, np.zeros(0, np.float32))
        _w_vec(out, rows.view(np.uint8).reshape(-1))
    else:
        out.append(struct.pack("<I", _fourcc({0: "IxFI", 1: "IxF2"}.get(metric, "IxFl"))))
        _w_header(out, d, n, metric)
        _w_vec(out, rows.reshape(-1))
    if id_map is not None:
        _w_vec(out, np.asarray(id_map, dtype=np.int64))
    with open(path, "wb") as f:
        f.write(b"".join(out))


def write_ivfflat(path, centroids, rows, assign, ids=None, nprobe=1, fp16=False, sparse_sizes=False):
    centroids = np.ascontiguousarray(centroids, dtype=np.float32)
    rows = np.ascontiguousarray(rows, dtype=np.float16 if fp16 else np.float32)
    nlist, d = centroids.shape
    n = rows.shape[0]
    ids = np.arange(n, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
    out = [struct.pack("<I", _fourcc("IwSq" if fp16 else "IwFl"))]
    _w_header(out, d, n, METRIC_INNER_PRODUCT)
    out.append(struct.pack("<QQ", nlist, nprobe))
    out.append(struct.pack("<I", _fourcc("IxFI")))
    _w_header(out, d, nlist, METRIC_INNER_PRODUCT)
    _w_vec(out, centroids.reshape(-1))
    out.append(struct.pack("<b", 0))
    _w_vec(out, np.zeros(0, np.int64))
    if fp16:
        out.append(struct.pack("<iifQQ", QT_FP16, 0, 0.0, d, 2 * d))
        _w_vec(out, np.zeros(0, np.float32))
        out.append(struct.pack("<QB", 2 * d, 0))
    else:
        out.append(struct.pack("<Q", 4 * d))
    out.append(struct.pack("<I", _fourcc("ilar")))
    out.append(struct.pack("<QQ", nlist, rows.dtype.itemsize * d))
    assign = np.asarray(assign, dtype=np.int64)
    sizes = np.bincount(assign, minlength=nlist).astype(np.uint64)
    if sparse_sizes:
        out.append(struct.pack("<I", _fourcc("sprs")))
        nz = np.nonzero(sizes)[0]
        _w_vec(out, np.stack([nz.astype(np.uint64), sizes[nz]], axis=1).reshape(-1))
    else:
        out.append(struct.pack("<I", _fourcc("full")))
        _w_vec(out, sizes)
    order = np.argsort(assign, kind="stable")
    off = 0
    for l in range(nlist):
        m = int(sizes[l])
        if m:
            sel = order[off:off + m]
            out.append(rows[sel].tobytes())
            out.append(ids[sel].tobytes())
            off += m
    with open(path, "wb") as f:
        f.write(b"".join(out))
