"""attention_tc2 softmax-loop variants (b200_attention_set_variant: bit 0 pipelined tcgen05.ld, bit 1 unmasked loop copies,
bit 2 P.V issued per 64-key block) at the benchmark shapes, one process, CUDA events over back-to-back launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_retrieval_b200._lib import lib, check

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
st = torch.cuda.current_stream().cuda_stream
old = lib.b200_attention_set_variant(-1)
VARIANTS = (0, 2, 6, 10, 14)   # + 'old' = the session-i build of the kernel (entry -4) as the fixed reference
for name, T, heads, causal in (("vision T=257", 257, 16, 0), ("vision T=256", 256, 16, 0), ("text T=77 causal", 77, 12, 1)):
    w = heads * 64
    qkv = torch.randn(B * T, 3 * w, device="cuda").bfloat16()
    out = torch.empty(B * T, w, device="cuda", dtype=torch.bfloat16)
    ref = None
    line = []
    for rnd in range(2):                      # two rounds: the order of the variants must not matter
        for v in VARIANTS:
            lib.b200_attention_set_variant(v)
            for _ in range(3):
                check(lib.b200_attention_tc_bf16_device(qkv.data_ptr(), None, -1, out.data_ptr(), B, T, heads, w, causal, 0, st), "attn")
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                check(lib.b200_attention_tc_bf16_device(qkv.data_ptr(), None, -1, out.data_ptr(), B, T, heads, w, causal, 0, st), "attn")
            e1.record()
            torch.cuda.synchronize()
            if ref is None:
                ref = out.clone()
            same = bool(torch.equal(ref, out))
            line.append("v%d %.3f%s" % (v, e0.elapsed_time(e1) / reps, "" if same else "(DIFF)"))
        for gen, nm in ((-4, "old"), (-2, "tc3")):
            for _ in range(3):
                check(lib.b200_attention_tc_bf16_device(qkv.data_ptr(), None, gen, out.data_ptr(), B, T, heads, w, causal, 0, st), "attn")
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                check(lib.b200_attention_tc_bf16_device(qkv.data_ptr(), None, gen, out.data_ptr(), B, T, heads, w, causal, 0, st), "attn")
            e1.record()
            torch.cuda.synchronize()
            line.append("%s %.3f%s" % (nm, e0.elapsed_time(e1) / 20, "" if gen != -4 or torch.equal(ref, out) else "(DIFF)"))
    print("%s B=%d ms: %s" % (name, B, "  ".join(line)), flush=True)
lib.b200_attention_set_variant(old)
