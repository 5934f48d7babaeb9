"""Target of the round's `ncu --set full` captures: one ViT-L/14 forward at batch 512 (pair GEMM, attention, LN), one
flat search at nq = 1 and nq = 1000 over 20M x 768 rows, one IVF search (nlist 16384, nprobe 16) at nq = 1000.
Select kernels with -k regex:... ; numbers printed by a run under ncu are never bench values."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clip_retrieval_b200 as m
from clip_retrieval_b200.index import synth_rows

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "embed"):
    arch = m.ARCHS["ViT-L/14"]
    model = m.B200Clip(arch, max_batch=512).load_state_dict(m.synthetic_state_dict(arch, seed=0))
    px = torch.randn(512, 3, 224, 224, device="cuda").clamp_(-1.8, 2.15)
    for _ in range(2):
        model.embed_image_device(px)
    torch.cuda.synchronize()
    del model
if what in ("all", "knn"):
    d = 768
    idx = m.B200FlatIndex(d)
    idx.reserve(20_000_000)
    for r0 in range(0, 20_000_000, 5_000_000):
        idx.add_synthetic(5_000_000, m.SynthSpec(seed=1234), row0=r0)
    q = synth_rows(1000, d, m.SynthSpec(seed=4321), dtype="float32")
    for _ in range(2):
        idx.search_device(q[:1].contiguous(), 40)
        idx.search_device(q, 40)
    torch.cuda.synchronize()
    del idx
if what in ("all", "ivf"):
    d, nlist = 768, 16384
    cent = synth_rows(nlist, d, m.SynthSpec(seed=7), dtype="float32").cpu().numpy()
    ivf = m.B200IVFFlatIndex(d, nlist, cent)
    ivf.add_synthetic(20_000_000, m.SynthSpec(seed=5, clustered=True, centroid_seed=7, nlist=nlist))
    ivf.nprobe = 16
    q = synth_rows(1000, d, m.SynthSpec(seed=77, clustered=True, centroid_seed=7, nlist=nlist), dtype="float32")
    for _ in range(2):
        ivf.search_device(q, 40)
    torch.cuda.synchronize()
