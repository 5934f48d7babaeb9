"""Generate tests/golden/knn_flat_768.npz: the float64 exhaustive ranking of a seeded synthetic set
(no FAISS offline).  Run from the repo root: python tests/golden/make_knn_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import knn_ref, synth_ref  # noqa: E402

n, d, nq, k, seed = 20000, 768, 8, 40, 1234
X = synth_ref.rows_f16(n, d, seed=seed)
for qseed in range(4321, 4400):  # first query seed whose top-(k+1) has no near-tie (ids arithmetic-independent)
    Q = synth_ref.rows_f32(nq, d, seed=qseed)
    S = knn_ref.scores_f64(X, Q)
    gaps = np.diff(np.sort(S, axis=1)[:, ::-1][:, :k + 1], axis=1)
    if np.abs(gaps).min() > 4e-6:
        break
else:
    raise SystemExit("no tie-free query seed found")
D, I = knn_ref.topk_from_scores(S, k)       # ranking in float64
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "knn_flat_768.npz"), D=D, I=I, n=n, seed=seed, qseed=qseed)
print("written; min score gap among the top-%d: %.3g" % (k + 1, np.abs(gaps).min()))
