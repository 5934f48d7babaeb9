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


class H14NsfwDetector:
    """Drop-in for `H14_NSFW_Detector` (clip_retrieval/h14_nsfw_model.py:10-35), the safety model
    `load_safety_model` returns for ViT-H/14 indices and `KnnService.get_unsafe_items` evaluates
    (clip_back.py:315-319): `predict(x, batch_size) -> np.float32 [n, 1]` logits.  The seven Linear
    layers run on the GPU in fp32 (C ABI `b200_mlp_*`); Dropout is the identity in eval, as in the
    reference.  Weights: `state_dict` in the reference's `nn.Sequential` key layout (`layers.0.weight`,
    `layers.0.bias`, `layers.3.weight`, ...), or `<cache_folder>/h14_nsfw_model/model.pt` — the file the
    reference downloads; there is no network here, so a missing file is an error."""

    DIMS = (1024, 2048, 1024, 256, 128, 16, 1)
    LINEAR_SLOTS = (0, 3, 6, 9, 12, 15, 16)   # positions of the nn.Linear modules inside the Sequential
    RELU = (1, 1, 1, 1, 1, 0, 0)

    def __init__(self, input_size=1024, cache_folder=None, state_dict=None, device=0):
        import os

        self.input_size = int(input_size)
        self.device_index = int(device)
        if state_dict is None:
            folder = cache_folder or os.path.expanduser("~/.cache/clip_retrieval")
            path = os.path.join(folder, "h14_nsfw_model", "model.pt")
            if not os.path.exists(path):
                raise FileNotFoundError("H14 NSFW weights not found at %s (the reference downloads h14_nsfw.pth there; "
                                        "no network here)" % path)
            state_dict = torch.load(path, map_location="cpu")
        dims = (self.input_size,) + self.DIMS
        self._h = C.c_void_p()
        cd = (C.c_int32 * len(dims))(*dims)
        cr = (C.c_uint8 * len(self.RELU))(*self.RELU)
        check(lib.b200_mlp_create(len(self.RELU), cd, cr, self.device_index, C.byref(self._h)), "mlp_create")
        for l, slot in enumerate(self.LINEAR_SLOTS):
            w = np.ascontiguousarray(_np(state_dict["layers.%d.weight" % slot]), dtype=np.float32)
            b = np.ascontiguousarray(_np(state_dict["layers.%d.bias" % slot]), dtype=np.float32)
            if w.shape != (dims[l + 1], dims[l]) or b.shape != (dims[l + 1],):
                raise ValueError("H14 NSFW layer %d: weight %r bias %r, expected %r" % (slot, w.shape, b.shape, (dims[l + 1], dims[l])))
            check(lib.b200_mlp_load_layer(self._h, l, C.c_void_p(w.ctypes.data), C.c_void_p(b.ctypes.data)), "mlp_load_layer")

    def __del__(self):
        try:
            if self._h:
                lib.b200_mlp_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def forward_device(self, x):
        """x: fp32 [n, input_size] (numpy or tensor) -> cuda fp32 [n, 1]."""
        E, dev = _rows_on_device(x, torch.device("cuda", self.device_index))
        if E.dim() != 2 or E.shape[1] != self.input_size:
            raise ValueError("H14 NSFW detector: expected [n, %d], got %r" % (self.input_size, tuple(E.shape)))
        y = torch.empty((E.shape[0], 1), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.b200_mlp_forward_device(self._h, C.c_void_p(E.data_ptr()), E.shape[0], C.c_void_p(y.data_ptr()),
                                          C.c_void_p(stream)), "mlp_forward")
        return y

    def predict(self, x, batch_size=None):
        """autokeras interface of the reference (h14_nsfw_model.py:43-48)."""
        del batch_size
        return self.forward_device(x).cpu().numpy()


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def get_unsafe_items(safety_model, embeddings, threshold=0.5):
    """Drop-in for KnnService.get_unsafe_items (clip_back.py:315-319)."""
    if hasattr(safety_model, "forward_device"):
        x = safety_model.forward_device(embeddings)[:, 0]
        return torch.nonzero(x > threshold).flatten().cpu().numpy()
    nsfw_values = safety_model.predict(embeddings, batch_size=embeddings.shape[0])
    x = np.array([e[0] for e in nsfw_values])
    return np.where(x > threshold)[0]
