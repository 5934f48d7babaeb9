"""TEST INFRASTRUCTURE ONLY — CPU oracle of the IVF coarse-quantiser training (csrc/knn_ivf.cu
`b200_kmeans_train_f16`).

The reference trains its indices through `autofaiss.build_index` (clip_retrieval/clip_index.py:12-31;
autofaiss>=2.17, faiss-cpu>=1.7.2 — requirements.txt:8,14, neither installable here), i.e. FAISS
`Clustering::train`: Lloyd iterations with the index's own assignment (IndexFlatIP: maximum inner
product), centroid = mean, `split_clusters` for empty clusters.  This restates that published algorithm
with the two choices the CUDA path makes explicit so that it is reproducible: the initial centroids are
one seeded pick per stride of the rows (FAISS: a random subset), and the cluster to split is the largest
one (FAISS: drawn with probability proportional to size).  Parity unpinned against FAISS itself.
Only tests/ may import this module."""
import numpy as np

MASK = (1 << 64) - 1


def _mix(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & MASK
    return x ^ (x >> 31)


def initial_picks(n, nlist, seed):
    out = []
    for i in range(nlist):
        lo, hi = (i * n) // nlist, ((i + 1) * n) // nlist
        span = max(1, hi - lo)
        out.append(lo + _mix(seed ^ ((i * 0xD1342543DE82EF95) & MASK)) % span)
    return np.array(out, np.int64)


def assign(X16, C32):
    """argmax inner product under fp16-rounded centroids, ties to the lower id (float64 scores)."""
    S = X16.astype(np.float64) @ C32.astype(np.float16).astype(np.float64).T
    return np.argmax(S, axis=1), S


def train(X16, nlist, niter, seed=1234, spherical=False):
    """Returns (centroids float32 [nlist, d], sizes int64 [nlist], assignment of the final centroids)."""
    n, d = X16.shape
    X32 = X16.astype(np.float32)
    C = X32[initial_picks(n, nlist, seed)].copy()
    eps = np.float32(1.0 / 1024.0)
    a = None
    for it in range(niter + 1):
        a, _ = assign(X16, C)
        sizes = np.bincount(a, minlength=nlist).astype(np.int64)
        if it == niter:
            break
        for l in range(nlist):
            rows = np.nonzero(a == l)[0]           # ascending row id: the order the device adds them in
            if len(rows) == 0:
                continue
            s = np.zeros(d, np.float32)
            for r in rows:
                s = s + X32[r]
            m = s * (np.float32(1.0) / np.float32(len(rows)))
            if spherical:
                nrm = np.sqrt(np.float32(np.sum(m.astype(np.float64) ** 2)))
                if nrm > 0:
                    m = m / nrm
            C[l] = m
        sz = sizes.copy()
        for ci in range(nlist):
            if sz[ci] != 0:
                continue
            cj = int(np.argmax(sz))                 # first maximum = lowest id among ties
            if sz[cj] < 2:
                break
            v = C[cj].copy()
            even = (np.arange(d) % 2) == 0
            C[ci] = np.where(even, v * (np.float32(1) + eps), v * (np.float32(1) - eps))
            C[cj] = np.where(even, v * (np.float32(1) - eps), v * (np.float32(1) + eps))
            sz[ci] = sz[cj] // 2
            sz[cj] -= sz[ci]
    return C, sizes, a
