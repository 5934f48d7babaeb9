"""Device post-filters (csrc/postfilter.cu) against the restated reference logic
(oracle/postfilter_ref.py <- clip_back.py:270-324).  Sets of dropped rows must be identical; the
synthetic rows keep every pair's inner product at least 1e-3 away from the threshold, so fp32
summation order cannot flip an edge."""
import numpy as np
import pytest
import torch

import clip_retrieval_b200 as b200
from oracle import postfilter_ref as R

pytestmark = pytest.mark.gpu


def _clustered(k, d, n_groups, seed, noise=0.02):
    """k unit rows: the first n_groups are 'originals', the rest near-copies of random originals
    (inner product ~0.999), plus chains a~b~c where a and c are NOT directly linked."""
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((n_groups, d)).astype(np.float32)
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    rows = [base]
    owner = rng.integers(0, n_groups // 2, k - n_groups)          # half of the originals stay unique
    dup = base[owner] + noise * rng.standard_normal((k - n_groups, d)).astype(np.float32) / np.sqrt(d)
    rows.append(dup / np.linalg.norm(dup, axis=1, keepdims=True))
    E = np.concatenate(rows).astype(np.float32)
    return E[rng.permutation(k)]


def _chain(d, length, step, seed):
    """Rows on a great circle, consecutive angle `step`: linked to their neighbours only."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(d); a /= np.linalg.norm(a)
    b = rng.standard_normal(d); b -= a * (a @ b); b /= np.linalg.norm(b)
    t = np.arange(length) * step
    return (np.cos(t)[:, None] * a + np.sin(t)[:, None] * b).astype(np.float32)


def _margin_ok(E, thr, eps=1e-3):
    S = E.astype(np.float64) @ E.astype(np.float64).T
    return np.abs(S - thr).min() > eps


@pytest.mark.parametrize("k,d", [(1, 768), (40, 768), (333, 512), (3000, 768), (4096, 64)])
def test_dedup_matches_reference_logic(k, d):
    E = _clustered(k, d, max(1, k // 3), seed=k) if k > 2 else _clustered(k, d, k, seed=1)
    assert _margin_ok(E, 0.94)
    want = R.get_non_uniques(E, 0.94)
    got = b200.get_non_uniques(E, 0.94)
    assert got == want
    if k >= 40:
        assert len(want) > 0


def test_dedup_follows_chains_and_keeps_lowest_index():
    d = 256
    chain = _chain(d, 12, 0.25, seed=3)             # cos(0.25)=0.969 > 0.94 > cos(0.5)=0.878: a path graph
    rng = np.random.default_rng(4)
    iso = rng.standard_normal((20, d)).astype(np.float32)
    iso /= np.linalg.norm(iso, axis=1, keepdims=True)
    E = np.concatenate([iso[:7], chain[::-1], iso[7:]])          # the chain's lowest index is its far end
    assert _margin_ok(E, 0.94)
    drop, labels = b200.dedup_mask(E, 0.94, return_labels=True)
    want = R.get_non_uniques(E, 0.94)
    assert torch.nonzero(drop).flatten().cpu().tolist() == want == list(range(8, 19))
    assert labels.cpu().tolist()[7:19] == [7] * 12


def test_dedup_ignores_nan_padding_rows():
    """Rows past the last result are NaN (clip_back.py:370-378 truncates them, but the filter must not choke)."""
    E = _clustered(64, 128, 20, seed=9)
    E[50:] = np.nan
    want = R.get_non_uniques(E[:50], 0.94)
    got = b200.get_non_uniques(E, 0.94)
    assert got == want


def test_dedup_accepts_device_rows_from_search():
    d, n, k = 128, 5000, 100
    rng = np.random.default_rng(0)
    X = rng.standard_normal((50, d)).astype(np.float32)
    X = np.repeat(X, n // 50, axis=0) + 0.01 * rng.standard_normal((n, d)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    idx = b200.B200FlatIndex(d)
    idx.add(X.astype(np.float16))
    q = torch.from_numpy(X[:1].copy()).cuda()
    D, I, Rr = idx.search_device(q, k, reconstruct=True)
    rows = Rr[0]
    got = b200.get_non_uniques(rows, 0.94)
    want = R.get_non_uniques(rows.cpu().numpy(), 0.94)
    assert got == want and len(got) > 50


def test_violence_detector_matches_einsum_argmax():
    rng = np.random.default_rng(2)
    E = rng.standard_normal((500, 768)).astype(np.float32)
    P = rng.standard_normal((3, 768)).astype(np.float32)
    np.testing.assert_array_equal(b200.get_violent_items(P, E), R.get_violent_items(P, E))


def test_dedup_rejects_oversized_k():
    with pytest.raises(b200.B200Error):
        b200.dedup_mask(np.zeros((4097, 8), np.float32))


@pytest.mark.timeout(120)
@pytest.mark.parametrize("n", [1, 40, 3000])
def test_h14_nsfw_head_matches_reference_module(n):
    """The H14 NSFW MLP (h14_nsfw_model.py:15-34, used at clip_back.py:315-319) on the GPU against the
    reference's own nn.Sequential stack instantiated with seeded weights (fp32 on the CPU)."""
    sd = R.h14_nsfw_state_dict(seed=3)
    rng = np.random.default_rng(n)
    E = (8.0 * rng.standard_normal((n, 1024))).astype(np.float32)   # spread the logits of the seeded weights
    det = b200.H14NsfwDetector(state_dict=sd)
    got = det.predict(E, batch_size=n)
    ref = R.h14_nsfw_predict(sd, E)
    assert got.shape == (n, 1) and got.dtype == np.float32
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=2e-5)
    thr = float(np.median(ref)) if n > 1 else 0.5     # seeded weights: put the decision boundary inside the data
    keep = np.nonzero(np.abs(ref[:, 0] - thr) > 1e-4)[0]   # rows not sitting on the threshold
    unsafe = b200.get_unsafe_items(det, E, threshold=thr)
    want = R.get_unsafe_items(sd, E, threshold=thr)
    assert set(np.intersect1d(unsafe, keep).tolist()) == set(np.intersect1d(want, keep).tolist())
    if n == 3000:
        assert 0.3 * n < len(want) < 0.7 * n          # the test exercises both outcomes
