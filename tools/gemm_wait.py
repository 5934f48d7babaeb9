"""Where does the MMA issuer of the CTA-pair GEMM wait?  Run with B200_GEMM_DEBUG=1 (in-kernel clock64 accounting)."""
import os
import sys

os.environ.setdefault("B200_GEMM_DEBUG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_retrieval_b200._lib import lib, check

M = int(sys.argv[1]) if len(sys.argv) > 1 else 131584
for name, N, K, act, res in (("qkv", 3072, 1024, 0, False), ("out", 1024, 1024, 0, True), ("fc", 4096, 1024, 1, False),
                             ("proj", 1024, 4096, 0, True), ("qkv-nobias", 3072, 1024, -1, False)):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    R = torch.randn(M, N, device="cuda", generator=g).bfloat16() if res else None
    Cc = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.b200_gemm_bf16_device(A.data_ptr(), W.data_ptr(), b.data_ptr() if act >= 0 else None,
                                        R.data_ptr() if res else None, Cc.data_ptr(), M, N, K, max(act, 0), 0, st), "gemm")
        e1.record()
        torch.cuda.synchronize()
    print("%s: %.3f ms incl. debug sync -> %.0f TFLOP/s" % (name, e0.elapsed_time(e1), 2.0 * M * N * K / e0.elapsed_time(e1) / 1e9), flush=True)
