"""GPU bring-up/profiling helper: the tcgen05 attention kernel alone (ViT-L/14 vision shape)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_retrieval_b200._lib import lib, check

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for (T, heads, causal) in ((257, 16, 0), (77, 12, 1)):
    w = heads * 64
    g = torch.Generator(device="cuda").manual_seed(1)
    qkv = torch.randn(B * T, 3 * w, device="cuda", generator=g).bfloat16()
    out = torch.empty(B * T, w, device="cuda", dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        check(lib.b200_attention_tc_bf16_device(qkv.data_ptr(), None, 0, out.data_ptr(), B, T, heads, w, causal, 0, st), "attn")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        lib.b200_attention_tc_bf16_device(qkv.data_ptr(), None, 0, out.data_ptr(), B, T, heads, w, causal, 0, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    items = B * heads
    tiles = items * ((T + 127) // 128)
    print("tc   T=%d B=%d heads=%d: %.3f ms  -> %.2f us per (b,h) per SM-slot, %.2f us per q-tile, %.1f TFLOP/s" % (
        T, B, heads, ms, ms * 1e3 / (items / 148), ms * 1e3 / (tiles / 148), 4.0 * T * T * 64 * items / ms / 1e9))
    for _ in range(2):
        lib.b200_attention_bf16_device(qkv.data_ptr(), out.data_ptr(), B, T, heads, w, causal, 0, st)
    e0.record()
    for _ in range(reps):
        lib.b200_attention_bf16_device(qkv.data_ptr(), out.data_ptr(), B, T, heads, w, causal, 0, st)
    e1.record()
    torch.cuda.synchronize()
    print("mma.sync variant: %.3f ms" % (e0.elapsed_time(e1) / reps))
    if T <= 264:
        for _ in range(2):
            lib.b200_attention_tc_bf16_device(qkv.data_ptr(), None, -1, out.data_ptr(), B, T, heads, w, causal, 0, st)
        e0.record()
        for _ in range(reps):
            lib.b200_attention_tc_bf16_device(qkv.data_ptr(), None, -1, out.data_ptr(), B, T, heads, w, causal, 0, st)
        e1.record()
        torch.cuda.synchronize()
        print("tc2 (two tiles in flight): %.3f ms" % (e0.elapsed_time(e1) / reps))
