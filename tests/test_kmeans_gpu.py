"""IVF coarse-quantiser training on the GPU (b200_kmeans_train_f16) against oracle/kmeans_ref.py.
With well-separated clusters every assignment is unambiguous, the device adds the rows of a list in
ascending row id like the oracle does, so the centroids must be bit-identical."""
import numpy as np
import pytest

import clip_retrieval_b200 as b200
from oracle import kmeans_ref as K
from oracle import knn_ref

pytestmark = pytest.mark.gpu


def _blobs(n, d, nc, seed, spread=0.15):
    """nc well-separated blobs, rows grouped blob after blob in equal shares: the trainer's one-pick-
    per-stride initialisation then starts with one centroid inside every blob, so no row is ever near
    a decision boundary and the device and the oracle must agree bit for bit."""
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((nc, d)).astype(np.float32)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    owner = (np.arange(n) * nc) // n
    x = c[owner] + spread * rng.standard_normal((n, d)).astype(np.float32) / np.sqrt(d)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float16)


@pytest.mark.parametrize("n,d,nlist,niter", [(4000, 64, 16, 5), (20000, 768, 64, 4), (3000, 128, 100, 3)])
def test_kmeans_matches_oracle(n, d, nlist, niter):
    X = _blobs(n, d, nlist, seed=n + d)
    cent, sizes = b200.train_kmeans(X, nlist, niter=niter, seed=7)
    want, wsizes, assign = K.train(X, nlist, niter, seed=7)
    # the comparison is only meaningful if no row sits on a decision boundary for the oracle
    _, S = K.assign(X, want)
    top2 = np.sort(S, axis=1)[:, -2:]
    assert (top2[:, 1] - top2[:, 0]).min() > 1e-4
    np.testing.assert_array_equal(sizes, wsizes)
    np.testing.assert_array_equal(cent, want)
    assert sizes.sum() == n


def test_kmeans_splits_empty_clusters():
    """More lists than distinct rows: the duplicates of a pick leave lists empty, which must be re-seeded
    from the largest cluster exactly as the oracle does."""
    d, nlist = 64, 8
    base = _blobs(3, d, 3, seed=1, spread=0.0)
    X = np.repeat(base, [50, 30, 20], axis=0)
    cent, sizes = b200.train_kmeans(X, nlist, niter=3, seed=3)
    want, wsizes, _ = K.train(X, nlist, 3, seed=3)
    np.testing.assert_array_equal(cent, want)
    np.testing.assert_array_equal(sizes, wsizes)


def test_build_ivf_index_from_trained_centroids():
    """train -> create -> add -> search: with nprobe = nlist the IVF result equals the exhaustive one, and
    the lists the index builds are the ones the trainer reported."""
    n, d, nlist, k = 30000, 256, 32, 10
    X = _blobs(n, d, nlist, seed=5)
    X = X[np.random.default_rng(0).permutation(n)]          # arbitrary order: lists of unequal sizes
    idx = b200.build_ivf_index(X, nlist, niter=4, seed=11, nprobe=nlist)
    cent, sizes = b200.train_kmeans(X, nlist, niter=4, seed=11)
    got_sizes, _ = idx.invlists()
    np.testing.assert_array_equal(np.asarray(got_sizes), sizes)
    Q = X[:16].astype(np.float32)
    D, I = idx.search(Q, k)
    ok, msg, _ = knn_ref.check_topk(D, I, knn_ref.scores_f64(X, Q), k, tol=2e-6)
    assert ok, msg
    idx.nprobe = 2
    D2, I2 = idx.search(Q, k)
    assert (I2[:, 0] == I[:, 0]).all()          # a row's own list is always probed first


def test_kmeans_rejects_bad_arguments():
    X = _blobs(100, 64, 4, seed=2)
    with pytest.raises(b200.B200Error):
        b200.train_kmeans(X, 200)               # more lists than rows
    with pytest.raises(b200.B200Error):
        b200.train_kmeans(X[:, :60], 4)         # d not a multiple of 8
