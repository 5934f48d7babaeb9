"""Throughput of the GPU image transform (csrc/preprocess.cu) vs the host transform it replaces
(torchvision on PIL images, one core): N synthetic 500x375 RGB images -> float32 [N,3,224,224]."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import clip_retrieval_b200 as b200
from clip_retrieval_b200.model import make_preprocess

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rng = np.random.default_rng(0)
imgs = [rng.integers(0, 256, (375, 500, 3), dtype=np.uint8) if i % 2 else rng.integers(0, 256, (500, 375, 3), dtype=np.uint8)
        for i in range(n)]
pre = b200.B200Preprocess(224)
buf, off, hh, ww = pre.pack(imgs)
pinned = torch.from_numpy(buf).pin_memory()
out = torch.empty((n, 3, 224, 224), dtype=torch.float32, device="cuda")
dev = pinned.cuda()
res = {}
for name, src in (("host_pixels", pinned.numpy()), ("device_pixels", dev)):
    for _ in range(3):
        pre.run_packed(src, off, hh, ww, out=out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        pre.run_packed(src, off, hh, ww, out=out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    res[name] = {"ms_per_batch": dt * 1e3, "images_per_s": n / dt, "pixel_GB_per_s": buf.nbytes / dt / 1e9}
from PIL import Image
tv = make_preprocess(224)
m = min(n, 128)
pil = [Image.fromarray(a) for a in imgs[:m]]
t0 = time.perf_counter()
ref = torch.stack([tv(im) for im in pil])
dt = time.perf_counter() - t0
res["torchvision_1core"] = {"images_per_s": m / dt}
res["bit_exact_vs_torchvision"] = bool(torch.equal(out[:m].cpu(), ref))
res["batch"] = n
print(json.dumps(res))
