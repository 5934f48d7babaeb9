"""Device post-filters of the search path: the reference's
`KnnService.get_non_uniques` and `get_violent_items` (clip_retrieval/clip_back.py:290-324) on the
reconstructed rows, which are already in HBM when `search_and_reconstruct` returns."""
import ctypes as C

import numpy as np
import torch

from ._lib import lib, check


def _rows_on_device(embeddings, device):
    if isinstance(embeddings, torch.Tensor):
        t = embeddings.detach()
    else:
        t = torch.from_numpy(np.ascontiguousarray(embeddings, dtype=np.float32))
    if device is None:
        device = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return t.to(device=device, dtype=torch.float32).contiguous(), torch.device(device)


def dedup_mask(embeddings, threshold=0.94, device=None, return_labels=False):
    """uint8 cuda tensor [k]: 1 for the rows `get_non_uniques` would return (clip_back.py:290-311)."""
    E, dev = _rows_on_device(embeddings, device)
    k, d = E.shape
    drop = torch.zeros(k, dtype=torch.uint8, device=dev)
    labels = torch.empty(k, dtype=torch.int32, device=dev)
    words = (k + 31) // 32
    ws = torch.empty(max(1, k * words), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    check(lib.b200_dedup_device(C.c_void_p(E.data_ptr()), k, d, C.c_float(threshold), C.c_void_p(drop.data_ptr()),
                                C.c_void_p(labels.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel() * 4),
                                dev.index or 0, C.c_void_p(stream)), "dedup")
    return (drop, labels) if return_labels else drop


def get_non_uniques(embeddings, threshold=0.94, device=None):
    """Drop-in for KnnService.get_non_uniques: list of row indices to remove (ascending)."""
    return torch.nonzero(dedup_mask(embeddings, threshold, device)).flatten().cpu().tolist()


def get_violent_items(safety_prompts, embeddings, device=None):
    """Drop-in for KnnService.get_violent_items (clip_back.py:321-324): np.where(argmax == 1)[0]."""
    E, dev = _rows_on_device(embeddings, device)
    P, _ = _rows_on_device(safety_prompts, dev)
    k, d = E.shape
    flag = torch.zeros(k, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    check(lib.b200_prompt_argmax_device(C.c_void_p(E.data_ptr()), k, d, C.c_void_p(P.data_ptr()), P.shape[0], 1,
                                        C.c_void_p(flag.data_ptr()), dev.index or 0, C.c_void_p(stream)), "prompt_argmax")
    return torch.nonzero(flag).flatten().cpu().numpy()
