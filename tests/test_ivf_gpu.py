"""IVF-Flat search path vs the CPU oracle (oracle/knn_ref.ivf_*): same centroids (fp16-rounded),
same assignment rule (max inner product), same nprobe -> identical ids, tie-aware in the scores."""
import numpy as np
import pytest

from oracle import knn_ref, synth_ref

pytestmark = pytest.mark.gpu
TOL = 2e-6


def _check(D, I, X, Q, assign, probes, k):
    for q in range(Q.shape[0]):
        rows = np.nonzero(np.isin(assign, probes[q][probes[q] >= 0]))[0].astype(np.int64)
        S = knn_ref.scores_f64(X[rows], Q[q:q + 1])
        ok, msg, _ = knn_ref.check_topk(D[q:q + 1], I[q:q + 1], S, k, ids=rows, tol=TOL)
        assert ok, "query %d: %s" % (q, msg)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("d,nlist,n", [(64, 16, 5000), (768, 37, 12000)])
def test_ivf_matches_oracle(d, nlist, n):
    import clip_retrieval_b200 as m

    X = synth_ref.rows_f16(n, d, seed=11, clustered=True, centroid_seed=7, nlist=nlist)
    C32 = synth_ref.centroids_f32(nlist, d, centroid_seed=7)
    C16 = C32.astype(np.float16)
    Q = synth_ref.rows_f32(6, d, seed=99, clustered=True, centroid_seed=7, nlist=nlist)
    idx = m.B200IVFFlatIndex(d, nlist, C32)
    idx.add(X[: n // 3])
    idx.finalize()
    idx.add(X[n // 3:].astype(np.float32))  # second batch: merged into the lists at the next finalize
    assert idx.ntotal == n and idx.nlist == nlist
    assign = knn_ref.ivf_assign(X, C16)
    sizes, ids = idx.invlists()
    assert np.array_equal(sizes, np.bincount(assign, minlength=nlist))
    off = 0
    for l in range(nlist):  # each list holds exactly the oracle's rows, in insertion order
        assert np.array_equal(ids[off:off + sizes[l]], np.nonzero(assign == l)[0])
        off += sizes[l]
    for nprobe in (1, 4, nlist):
        idx.nprobe = nprobe
        assert idx.nprobe == nprobe
        for k in (1, 40):
            D, I, R = idx.search_and_reconstruct(Q, k)
            Do, Io, probes = knn_ref.ivf_search(X, assign, C16, Q, k, nprobe)
            _check(D, I, X, Q, assign, probes, k)
            assert np.array_equal(I, Io)  # no near-ties in this seeded set: ids identical
            np.testing.assert_allclose(D, Do, atol=TOL)
            assert np.array_equal(R[I >= 0], X[I[I >= 0]].astype(np.float32))
    # probing every list is the exhaustive search
    idx.nprobe = nlist
    D, I = idx.search(Q, 40)
    Df, If = knn_ref.flat_search(X, Q, 40)
    assert np.array_equal(I, If)


@pytest.mark.timeout(300)
def test_ivf_synthetic_by_construction_and_small_lists():
    import clip_retrieval_b200 as m

    d, nlist, n, k = 768, 64, 40000, 40
    spec = m.SynthSpec(seed=5, clustered=True, centroid_seed=7, nlist=nlist, cw=3, nw=1)
    C32 = synth_ref.centroids_f32(nlist, d, centroid_seed=7)
    idx = m.B200IVFFlatIndex(d, nlist, C32)
    idx.add_synthetic(n, spec)  # bucketed by generating list, generated straight into list order
    X = synth_ref.rows_f16(n, d, seed=5, clustered=True, centroid_seed=7, nlist=nlist)
    assign = synth_ref.list_of_rows(7, np.arange(n), nlist)
    sizes, ids = idx.invlists()
    assert np.array_equal(sizes, np.bincount(assign, minlength=nlist))
    Q = synth_ref.rows_f32(9, d, seed=77, clustered=True, centroid_seed=7, nlist=nlist)
    idx.nprobe = 8
    idx.id_base = 1_000_000
    D, I, R = idx.search_and_reconstruct(Q, k)
    Do, Io, probes = knn_ref.ivf_search(X, assign, C32.astype(np.float16), Q, k, 8, id_base=1_000_000)
    assert np.array_equal(I, Io)
    np.testing.assert_allclose(D, Do, atol=TOL)
    assert np.array_equal(R, X[I - 1_000_000].astype(np.float32))
    ms, launches = idx.last_scan_ms()
    assert launches == 1 and ms > 0
    # k larger than the probed lists hold: -1 padding
    idx.nprobe = 1
    D, I = idx.search(Q[:2], 2000)
    held = np.array([sizes[p] for p in probes[:2, 0]])
    for q in range(2):
        assert (I[q, :held[q]] >= 0).all() and (I[q, held[q]:] == -1).all()


def test_ivf_range_search_and_reconstruct_by_id():
    """index.range_search / index.reconstruct on an IVF index (clip_filter.py:52 calls range_search on whatever
    index was loaded): hits = rows of the probed lists above the threshold; reconstruct(id) finds the row through
    the id -> slot map of the list-ordered store."""
    import clip_retrieval_b200 as m
    from oracle import knn_ref, synth_ref

    d, n, nlist = 128, 30000, 32
    kw = dict(seed=5, clustered=True, centroid_seed=7, nlist=nlist, cw=3, nw=1)
    X = synth_ref.rows_f16(n, d, **kw)
    C = synth_ref.centroids_f32(nlist, d, 7)
    C16 = C.astype(np.float16)
    idx = m.B200IVFFlatIndex(d, nlist, C)
    idx.add(X)
    idx.id_base = 1000
    Q = synth_ref.rows_f32(3, d, seed=77, clustered=True, centroid_seed=7, nlist=nlist)
    assign = knn_ref.ivf_assign(X, C16)
    for nprobe, thr in ((1, 0.6), (4, 0.3), (nlist, 0.5)):
        idx.nprobe = nprobe
        lims, D, I = idx.range_search(Q, thr)
        _, probes = knn_ref.topk_from_scores(knn_ref.scores_f32(C16, Q), nprobe)
        S64 = knn_ref.scores_f64(X, Q)
        for q in range(3):
            rows = np.nonzero(np.isin(assign, probes[q]))[0]
            want = set((rows[S64[q, rows] > thr] + 1000).tolist())
            got = I[lims[q]:lims[q + 1]]
            assert np.all(np.diff(got) > 0)
            for i in set(got.tolist()) ^ want:
                assert abs(S64[q, i - 1000] - thr) <= 2e-6, "id %d differs and is not at the threshold" % i
            np.testing.assert_allclose(D[lims[q]:lims[q + 1]], S64[q, got - 1000], atol=2e-6)
        assert lims[-1] > 0
    for i in (0, 1, n // 2, n - 1):
        assert np.array_equal(idx.reconstruct(1000 + i), X[i].astype(np.float32))
    assert np.isnan(idx.reconstruct(-1)).all() and np.isnan(idx.reconstruct(1000 + n)).all()


def test_load_index_reads_faiss_files(tmp_path):
    """load_index on `image.index` files in the FAISS layout (clip_back.py:589-596): flat fp32, fp16 with an id map,
    IVF-Flat whose inverted lists are NOT the max-inner-product assignment (the file's own lists must be kept)."""
    import clip_retrieval_b200 as m
    from clip_retrieval_b200 import faiss_io
    from oracle import knn_c, knn_ref, synth_ref

    d, n, nlist, k = 64, 5000, 16, 20
    X = synth_ref.rows_f16(n, d)
    Q = synth_ref.rows_f32(5, d, seed=4321)
    p = str(tmp_path / "image.index")
    faiss_io.write_flat(p, X.astype(np.float32))
    idx = m.load_index(p)
    D, I = idx.search(Q, k)
    ok, msg, _ = knn_ref.check_topk(D, I, knn_ref.scores_f64(X, Q), k)
    assert ok, msg
    ids = np.arange(n, dtype=np.int64) * 3 + 11
    faiss_io.write_flat(p, X, fp16=True, id_map=ids)
    idx = m.load_index(p, enable_faiss_memory_mapping=True)
    D2, I2, R2 = idx.search_and_reconstruct(Q, k)
    assert np.array_equal(I2, ids[I]) and np.array_equal(D2, D) and np.array_equal(R2, knn_ref.reconstruct(X, I))
    # IVF: lists by a rule that is not the argmax (row id mod nlist); the loaded index must scan exactly those lists
    C = synth_ref.centroids_f32(nlist, d, 7)
    assign = np.arange(n) % nlist
    folder = tmp_path / "populated_dir"
    folder.mkdir()
    faiss_io.write_ivfflat(str(folder / "populated.index"), C, X, assign, nprobe=3, fp16=True)
    ivf = m.load_index(str(folder))
    assert ivf.nprobe == 3 and ivf.ntotal == n and ivf.nlist == nlist
    Dv, Iv = ivf.search(Q, k)
    Xl, off, lids = knn_c.ivf_layout(X, assign, nlist)
    Do, Io, _ = knn_c.ivf_search(Xl, off, lids, C.astype(np.float16), Q, k, 3)
    assert np.array_equal(Iv, Io)
    np.testing.assert_allclose(Dv, Do, atol=2e-6)
    with pytest.raises(NotImplementedError):
        open(p, "wb").write(b"IwPQ" + b"\0" * 64)
        m.load_index(p)
