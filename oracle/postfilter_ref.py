"""TEST INFRASTRUCTURE ONLY — CPU oracle of the search post-filters.

Restates `KnnService.connected_components`, `get_non_uniques` and `get_violent_items`
(reference clip_retrieval/clip_back.py:270-324) with FAISS `IndexFlatIP.range_search(x, thr)` written
out as what it computes: all (i, j) whose fp32 inner product is strictly greater than `thr`
(faiss-cpu >= 1.7.2, requirements.txt:8 — not installable here; parity unpinned against FAISS itself,
the reference's tests hold no vector for this path).  Only tests/ may import this module."""
from collections import defaultdict

import numpy as np


def range_pairs(embeddings, threshold):
    S = embeddings.astype(np.float32) @ embeddings.astype(np.float32).T
    return S > np.float32(threshold)


def connected_components(neighbors):
    """clip_back.py:270-288, verbatim semantics (iteration in insertion order, first node leads)."""
    seen = set()

    def component(node):
        r = []
        nodes = set([node])
        while nodes:
            node = nodes.pop()
            seen.add(node)
            nodes |= set(neighbors[node]) - seen
            r.append(node)
        return r

    u = []
    for node in list(neighbors):
        if node not in seen:
            u.append(component(node))
    return u


def get_non_uniques(embeddings, threshold=0.94, adjacency=None):
    """clip_back.py:290-311; returns the sorted list of indices the reference would drop."""
    A = range_pairs(embeddings, threshold) if adjacency is None else adjacency
    same_mapping = defaultdict(list)
    for i in range(A.shape[0]):
        for j in np.nonzero(A[i])[0]:
            same_mapping[int(i)].append(int(j))
    non_uniques = set()
    for g in connected_components(same_mapping):
        for e in g[1:]:
            non_uniques.add(e)
    return sorted(non_uniques)


def get_violent_items(safety_prompts, embeddings):
    """clip_back.py:321-324."""
    pred = np.einsum("ij,kj->ik", embeddings, safety_prompts)
    return np.where(np.argmax(pred, axis=1) == 1)[0]


def h14_nsfw_layers(input_size=1024):
    """The module stack of `H14_NSFW_Detector.__init__` (clip_retrieval/h14_nsfw_model.py:15-34) restated
    with torch.nn (weights are downloaded by the reference; tests seed them)."""
    from torch import nn

    return nn.Sequential(
        nn.Linear(input_size, 1024), nn.ReLU(), nn.Dropout(0.2),
        nn.Linear(1024, 2048), nn.ReLU(), nn.Dropout(0.2),
        nn.Linear(2048, 1024), nn.ReLU(), nn.Dropout(0.2),
        nn.Linear(1024, 256), nn.ReLU(), nn.Dropout(0.2),
        nn.Linear(256, 128), nn.ReLU(), nn.Dropout(0.2),
        nn.Linear(128, 16),
        nn.Linear(16, 1),
    )


def h14_nsfw_state_dict(seed=0, input_size=1024):
    import torch

    torch.manual_seed(seed)
    net = h14_nsfw_layers(input_size)
    return {"layers." + k: v.detach().clone() for k, v in net.state_dict().items()}


def h14_nsfw_predict(state_dict, x):
    """`H14_NSFW_Detector.predict` (h14_nsfw_model.py:43-48): eval-mode forward in fp32 on the CPU."""
    import torch

    net = h14_nsfw_layers(x.shape[1]).eval()
    net.load_state_dict({k[len("layers."):]: v for k, v in state_dict.items()})
    with torch.no_grad():
        return net(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))).numpy()


def get_unsafe_items(state_dict, embeddings, threshold=0.5):
    """clip_back.py:315-319."""
    x = np.array([e[0] for e in h14_nsfw_predict(state_dict, embeddings)])
    return np.where(x > threshold)[0]
