"""GPU image transform: the `preprocess` step of the reference's readers
(clip_retrieval/clip_inference/reader.py:98-106,158-165) moved behind the C ABI.

`B200Preprocess(n_px)(images)` takes decoded images (PIL.Image or uint8 HWC arrays, any sizes) and
returns the float32 `[n, 3, n_px, n_px]` batch on the device that `ClipMapper` / `encode_image`
consume — bit-identical to torchvision's Compose[Resize(bicubic), CenterCrop, ToTensor, Normalize]
on the host (tests/test_preprocess_gpu.py; pinned on the reference's own test_tensors fixtures).
JPEG decoding stays where the reference does it (PIL in the DataLoader workers); what crosses PCIe is
the decoded uint8 image instead of the resized float32 tensor."""
import ctypes as C

import numpy as np
import torch

from ._lib import lib, check

OPENAI_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_STD = (0.26862954, 0.26130258, 0.27577711)


def to_rgb8(image):
    """PIL.Image | ndarray -> contiguous uint8 [H, W, 3].  Mode conversion happens before the resize
    (the reference converts after it; identical for RGB and L sources, which is what JPEG yields)."""
    if isinstance(image, np.ndarray):
        a = image
    else:
        a = np.asarray(image.convert("RGB"))
    if a.ndim == 2:
        a = np.repeat(a[:, :, None], 3, axis=2)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise ValueError(f"expected uint8 [H,W,3] pixels, got {a.dtype} {a.shape}")
    return np.ascontiguousarray(a)


def pack_images(images):
    """-> (uint8 buffer, int64 offsets, int32 heights, int32 widths): the C-ABI batch layout of
    b200_preproc_run (images back to back, RGB uint8 HWC)."""
    arrs = [to_rgb8(im) for im in images]
    heights = np.array([a.shape[0] for a in arrs], np.int32)
    widths = np.array([a.shape[1] for a in arrs], np.int32)
    sizes = heights.astype(np.int64) * widths * 3
    offsets = np.zeros(len(arrs), np.int64)
    if len(arrs) > 1:
        offsets[1:] = np.cumsum(sizes[:-1])
    buf = np.empty(int(sizes.sum()), np.uint8)
    for a, o, s in zip(arrs, offsets, sizes):
        buf[o:o + s] = a.reshape(-1)
    return buf, offsets, heights, widths


class B200Preprocess:
    def __init__(self, n_px=224, mean=OPENAI_MEAN, std=OPENAI_STD, device=None):
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.n_px = int(n_px)
        m = (C.c_float * 3)(*mean)
        s = (C.c_float * 3)(*std)
        h = C.c_void_p()
        check(lib.b200_preproc_create(self.n_px, m, s, self.device.index or 0, C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.b200_preproc_destroy(h)

    def pack(self, images):
        return pack_images(images)

    def run_packed(self, pixels, offsets, heights, widths, out=None):
        """pixels: uint8 numpy buffer (host) or uint8 cuda tensor; returns float32 [n,3,n_px,n_px]."""
        n = len(heights)
        if out is None:
            out = torch.empty((n, 3, self.n_px, self.n_px), dtype=torch.float32, device=self.device)
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.shape[0] == n
        on_dev = isinstance(pixels, torch.Tensor) and pixels.is_cuda
        ptr = pixels.data_ptr() if isinstance(pixels, torch.Tensor) else pixels.ctypes.data
        offsets = np.ascontiguousarray(offsets, np.int64)
        heights = np.ascontiguousarray(heights, np.int32)
        widths = np.ascontiguousarray(widths, np.int32)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib.b200_preproc_run(self._h, C.c_void_p(ptr), int(on_dev), C.c_void_p(offsets.ctypes.data),
                                   C.c_void_p(heights.ctypes.data), C.c_void_p(widths.ctypes.data), n,
                                   C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
        return out

    def __call__(self, images, out=None):
        if not isinstance(images, (list, tuple)):
            images = [images]
        return self.run_packed(*self.pack(images), out=out)
