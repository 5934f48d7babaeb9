"""Attention kernel timing at the benchmark shapes: ViT-L/14 vision (B x 257 tokens, 16 heads) and text (B x 77, 12 heads,
causal), generation 2 (attention_tc2) vs 3 (attention_tc3), CUDA events over back-to-back launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_retrieval_b200._lib import lib, check

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
st = torch.cuda.current_stream().cuda_stream
shapes = [("vision", 257, 16, 0), ("text", 77, 12, 1)]
if len(sys.argv) > 2:   # extra token counts for the vision shape, e.g. 256 (no leftover row), 129, 260
    shapes = [("vision T=%s" % t, int(t), 16, 0) for t in sys.argv[2:]]
for name, T, heads, causal in shapes:
    w = heads * 64
    qkv = torch.randn(B * T, 3 * w, device="cuda").bfloat16()
    out = torch.empty(B * T, w, device="cuda", dtype=torch.bfloat16)
    res = {}
    for gen in (-1, -3, -2):
        for _ in range(3):
            check(lib.b200_attention_tc_bf16_device(qkv.data_ptr(), None, gen, out.data_ptr(), B, T, heads, w, causal, 0, st), "attn")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            check(lib.b200_attention_tc_bf16_device(qkv.data_ptr(), None, gen, out.data_ptr(), B, T, heads, w, causal, 0, st), "attn")
        e1.record()
        torch.cuda.synchronize()
        res[gen] = e0.elapsed_time(e1) / reps
        flops = 4.0 * T * T * w * B
        print("%s B=%d gen %d: %.3f ms  (%.0f TFLOP/s, %.0f clk/head at 1.9 GHz x 148 SMs)" % (
            name, B, 1 - gen, res[gen], flops / res[gen] / 1e9, res[gen] * 1e-3 * 1.9e9 * 148 / (B * heads)), flush=True)
