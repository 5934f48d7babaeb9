"""Does keeping a layer's activations inside the 126 MB L2 pay?  ViT-L/14 at 1024 pairs per step, with the towers run
over sub-batches of `max_batch` samples (the handle already loops over chunks of max_batch): smaller chunks keep
x / qkv / attention output / MLP hidden of a chunk L2-resident between the GEMMs of a layer (less HBM traffic, less
power) at the price of more tile-wave quantisation and launches.  Prints pairs/s per chunk size."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clip_retrieval_b200 as m

arch = m.ARCHS["ViT-L/14"]
sd = m.synthetic_state_dict(arch, seed=0)
B = 1024
g = torch.Generator().manual_seed(0)
px = torch.randn(B, 3, 224, 224, generator=g).clamp_(-1.8, 2.15).cuda()
tok = torch.zeros(B, 77, dtype=torch.int64)
tok[:, 0], tok[:, 1:20], tok[:, 20] = arch.vocab_size - 2, 1000, arch.vocab_size - 1
tok = tok.cuda()
for mb in [int(x) for x in (sys.argv[1:] or ["1024", "512", "256", "128", "64"])]:
    model = m.B200Clip(arch, max_batch=mb).load_state_dict(sd)
    for _ in range(2):
        model.embed_image_device(px); model.embed_text_device(tok)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 5
    for _ in range(K):
        model.embed_image_device(px); model.embed_text_device(tok)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    model.set_profiling(True)
    model.embed_image_device(px); model.embed_text_device(tok)
    torch.cuda.synchronize()
    tm = model.last_timing()
    model.set_profiling(False)
    print("max_batch %4d: %.1f ms/step = %.0f pairs/s | gemm %.1f attn %.1f ln %.1f other %.1f | by kind %s" % (
        mb, ms, B / ms * 1e3, tm["gemm"], tm["attention"], tm["layernorm"], tm["other"],
        {k: round(v, 1) for k, v in tm["gemm_by_kind"].items()}), flush=True)
    del model
    torch.cuda.empty_cache()
