"""faiss_io: the FAISS index-file layouts `load_index` accepts (clip_back.py:589-596 calls faiss.read_index).  FAISS is
not installable here, so these are round trips through this module's own writer of the published layout — the reader
is UNVERIFIED against FAISS-written files (stated in faiss_io.py and DESIGN.md)."""
import os
import struct

import numpy as np
import pytest

from clip_retrieval_b200 import faiss_io
from oracle import synth_ref


def test_flat_layouts_round_trip(tmp_path):
    d = 48
    X = synth_ref.rows_f16(257, d)
    p = str(tmp_path / "image.index")
    faiss_io.write_flat(p, X.astype(np.float32))
    r = faiss_io.read_faiss_index(p)
    assert r["kind"] == "flat" and r["d"] == d and r["ntotal"] == 257 and r["ids"] is None
    assert r["rows"].dtype == np.float32 and np.array_equal(r["rows"], X.astype(np.float32))
    # byte layout of the header: fourcc, d, ntotal, two dummies, is_trained, metric
    raw = open(p, "rb").read()
    assert raw[:4] == b"IxFI" and struct.unpack_from("<iq", raw, 4) == (d, 257) and struct.unpack_from("<Bi", raw, 32) == (1, 0)
    assert struct.unpack_from("<Q", raw, 37)[0] == 257 * d           # vector length in 4-byte units
    faiss_io.write_flat(p, X, fp16=True)
    r = faiss_io.read_faiss_index(p)
    assert r["rows"].dtype == np.float16 and np.array_equal(r["rows"], X)
    ids = np.arange(257, dtype=np.int64)[::-1] * 7 + 3
    faiss_io.write_flat(p, X, fp16=True, id_map=ids)
    r = faiss_io.read_faiss_index(p)
    assert np.array_equal(r["ids"], ids) and np.array_equal(r["rows"], X)
    assert faiss_io.looks_like_faiss_index(p) and not faiss_io.looks_like_faiss_index(__file__)


@pytest.mark.parametrize("fp16,sparse", [(False, False), (True, True)])
def test_ivf_layouts_round_trip(tmp_path, fp16, sparse):
    d, nlist, n = 32, 11, 400
    X = synth_ref.rows_f16(n, d)
    C = synth_ref.centroids_f32(nlist, d, 7)
    assign = (np.arange(n) * 7) % 5 if sparse else (np.arange(n) * 7) % nlist   # sparse: lists 5..10 stay empty
    ids = np.arange(n, dtype=np.int64) * 2 + 1
    p = str(tmp_path / "text.index")
    faiss_io.write_ivfflat(p, C, X, assign, ids=ids, nprobe=3, fp16=fp16, sparse_sizes=sparse)
    r = faiss_io.read_faiss_index(p)
    assert r["kind"] == "ivfflat" and r["nlist"] == nlist and r["nprobe"] == 3 and r["ntotal"] == n
    assert np.array_equal(r["centroids"], C) and np.array_equal(r["list_sizes"], np.bincount(assign, minlength=nlist))
    order = np.argsort(assign, kind="stable")
    assert np.array_equal(r["ids"], ids[order]) and np.array_equal(np.asarray(r["rows"], np.float16), X[order])


def test_unsupported_types_name_the_way_out(tmp_path):
    p = str(tmp_path / "image.index")
    for tag in (b"IwPQ", b"IxPT", b"IHNf", b"IxPq"):
        open(p, "wb").write(tag + b"\0" * 64)
        with pytest.raises(NotImplementedError, match="load_index"):
            faiss_io.read_faiss_index(p)
    faiss_io.write_flat(p, np.zeros((3, 8), np.float32), metric=faiss_io.METRIC_L2)
    with pytest.raises(NotImplementedError, match="inner product"):
        faiss_io.read_faiss_index(p)
