"""Timing experiment for the next GEMM design step: how fast do the model's four GEMM shapes run when
each SM pulls 48 instead of 64 operand bytes per clock from L2 (B200_GEMM_HALFB=1: each CTA of a pair
loads only half of its B rows — the traffic of a 2-pair cluster with a multicast B tile; results are
numerically wrong in that mode, only the time is meaningful)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_retrieval_b200._lib import lib

M = 1024 * 257
st = torch.cuda.current_stream().cuda_stream
for name, N, K in (("qkv", 3072, 1024), ("out", 1024, 1024), ("fc", 4096, 1024), ("c_proj", 1024, 4096)):
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    Cc = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        lib.b200_gemm_bf16_device(A.data_ptr(), W.data_ptr(), None, None, Cc.data_ptr(), M, N, K, 0, 0, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 30
    e0.record()
    for _ in range(reps):
        lib.b200_gemm_bf16_device(A.data_ptr(), W.data_ptr(), None, None, Cc.data_ptr(), M, N, K, 0, 0, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%s %-7s M=%d N=%d K=%d: %.3f ms  %.0f TFLOP/s" % ("HALFB" if os.environ.get("B200_GEMM_HALFB") else "full ", name, M, N, K, ms,
                                                              2.0 * M * N * K / ms / 1e9))
    del A, W, Cc
