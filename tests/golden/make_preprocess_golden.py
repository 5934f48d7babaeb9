"""Generate tests/golden/preprocess_ref.npz from the reference's own preprocess fixtures.

The reference holds, for its 7 test images, the tensors its `preprocess` produced
(/root/reference/tests/test_clip_inference/test_images/*.jpg ->
 /root/reference/tests/test_clip_inference/test_tensors/*.pkl, written by playground.ipynb cell 9
 with the OpenAI ViT-B/32 transform).  This script (run in the build container, where /root/reference
exists) decodes the JPEGs with Pillow, checks that the CPU oracle reproduces the reference tensors bit
for bit, and commits: the decoded uint8 pixels of four images (both orientations, small and large), and
for all seven the sha256 of the reference's float32 tensor, so that the GPU test can verify the CUDA
transform against the reference's own golden vectors without /root/reference.

    python tests/golden/make_preprocess_golden.py
"""
import glob
import hashlib
import os
import pickle
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import preprocess_ref as P  # noqa: E402

REF = "/root/reference/tests/test_clip_inference"
KEEP_PIXELS = ["123_456", "456_123", "416_264", "321_421"]


def main():
    out = {}
    names = []
    for f in sorted(glob.glob(REF + "/test_tensors/*.pkl")):
        with open(f, "rb") as fh:
            o = pickle.load(fh)
        for name, t in zip(o["image_filename"], o["image_tensor"]):
            ref = np.ascontiguousarray(t.numpy())
            px = np.asarray(Image.open(f"{REF}/test_images/{name}.jpg").convert("RGB"))
            got = P.preprocess(px)
            assert got.dtype == ref.dtype and np.array_equal(got, ref), name   # oracle == reference fixture
            names.append(name)
            out[f"sha256_{name}"] = np.frombuffer(hashlib.sha256(ref.tobytes()).digest(), np.uint8)
            out[f"shape_{name}"] = np.array(px.shape[:2], np.int32)
            if name in KEEP_PIXELS:
                out[f"pixels_{name}"] = px
    out["names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "preprocess_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(names), "images pinned")


if __name__ == "__main__":
    main()
