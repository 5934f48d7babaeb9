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
