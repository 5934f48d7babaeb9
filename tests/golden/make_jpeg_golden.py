"""Generate tests/golden/jpeg/*.jpg + jpeg_pixels.npz: small synthetic JPEG files written with Pillow (baseline,
4:4:4 and 4:2:0 chroma, one grayscale) and the pixels Pillow/libjpeg-turbo decodes from them — the decode the
reference's readers perform (reader.py:98-106).  The GPU decode (nvJPEG) is compared against these on the GPU box.
    python tests/golden/make_jpeg_golden.py
"""
import io
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))


def scene(h, w, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([
        128 + 100 * np.sin(x / 17.0 + seed) * np.cos(y / 23.0),
        128 + 90 * np.cos((x + y) / 31.0),
        128 + 80 * np.sin(y / 13.0 - x / 41.0),
    ], axis=2)
    img += rng.normal(0, 6, img.shape)
    cy, cx = h // 3, w // 2
    img[cy:cy + h // 4, cx:cx + w // 5] = (220, 40, 60)          # a hard-edged block (chroma upsampling differs there)
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    out = {}
    for name, (h, w, seed, kw) in {
        "a_444": (180, 240, 1, dict(quality=92, subsampling=0)),
        "b_420": (333, 250, 2, dict(quality=85, subsampling=2)),
        "c_gray": (120, 97, 3, dict(quality=90)),
    }.items():
        arr = scene(h, w, seed)
        im = Image.fromarray(arr if name != "c_gray" else arr[:, :, 0])
        buf = io.BytesIO()
        im.save(buf, format="JPEG", **kw)
        data = buf.getvalue()
        open(os.path.join(HERE, "jpeg", name + ".jpg"), "wb").write(data)
        out[name] = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        print(name, len(data), "bytes", out[name].shape)
    np.savez_compressed(os.path.join(HERE, "jpeg_pixels.npz"), **out)


if __name__ == "__main__":
    main()
