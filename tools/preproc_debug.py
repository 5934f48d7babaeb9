"""Bring-up helper: run the GPU image transform on a few ragged images and compare with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import clip_retrieval_b200 as b200
from oracle import preprocess_ref as P

shapes = [(123, 456), (456, 123), (416, 264), (321, 421), (224, 224), (1, 1)]
imgs = [np.random.default_rng(i).integers(0, 256, (*s, 3), dtype=np.uint8) for i, s in enumerate(shapes)]
pre = b200.B200Preprocess(224)
for i, im in enumerate(imgs):
    out = pre([im]).cpu().numpy()[0]
    want = P.preprocess(im)
    print(shapes[i], np.array_equal(out, want), float(np.abs(out - want).max()))
