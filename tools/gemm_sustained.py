# SPEC_MODES: --spec-ab times the run-time epilogue against the compile-time flavours
"""Sustained (power-capped) throughput of the model's four GEMM shapes with and without their epilogue
features: each configuration runs back to back for ~1.5 s after a 1 s warm-up of the same kernel, so the
clocks are the ones the model sees, not burst clocks."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_retrieval_b200._lib import lib

M = 1024 * 257
st = torch.cuda.current_stream().cuda_stream


def run(name, N, K, bias, res, act, secs=0.6):
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda") if bias else None
    R = torch.randn(M, N, device="cuda").bfloat16() if res else None
    Cc = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

    def go():
        lib.b200_gemm_bf16_device(A.data_ptr(), W.data_ptr(), b.data_ptr() if bias else None, R.data_ptr() if res else None,
                                  Cc.data_ptr(), M, N, K, act, 0, st)
    t0 = time.time()
    while time.time() - t0 < 0.6:
        for _ in range(20):
            go()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(20, int(secs / 0.0015))
    e0.record()
    for _ in range(reps):
        go()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-7s N=%d K=%d bias=%d res=%d act=%d: %.3f ms  %.0f TFLOP/s" % (name, N, K, bias, res, act, ms, 2.0 * M * N * K / ms / 1e9),
          flush=True)


def run_cublas(name, N, K, secs=0.6):
    """The vendor library on the same shape (torch.nn.functional.linear -> cuBLASLt, bias epilogue), same sustained protocol:
    a per-shape reference point for the hand-written kernel — MEASURED_PEAKS' sustained figure is cuBLAS at a large square shape."""
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16()
    t0 = time.time()
    while time.time() - t0 < 0.6:
        for _ in range(20):
            torch.nn.functional.linear(A, W, b)
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(20, int(secs / 0.0015))
    e0.record()
    for _ in range(reps):
        torch.nn.functional.linear(A, W, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-7s N=%d K=%d cuBLASLt linear+bias:      %.3f ms  %.0f TFLOP/s" % (name, N, K, ms, 2.0 * M * N * K / ms / 1e9), flush=True)


# the four per-layer GEMMs with the epilogue features the model uses: (bias, residual, activation)
for name, N, K, feats in (("qkv", 3072, 1024, [(1, 0, 0)]),
                          ("out", 1024, 1024, [(1, 0, 0), (1, 1, 0)]),
                          ("fc", 4096, 1024, [(1, 0, 0), (1, 0, 1)]),
                          ("c_proj", 1024, 4096, [(1, 1, 0)])):
    for bias, res, act in feats:
        for mode in ((2, 1) if "--spec-ab" in sys.argv else (1,)):   # 2 = run-time epilogue, 1 = compile-time flavour (default)
            lib.b200_gemm_set_tma_store(mode)
            if "--spec-ab" in sys.argv:
                print("  [epilogue %s]" % ("run-time" if mode == 2 else "compile-time"), end=" ")
            run(name, N, K, bias, res, act)
        lib.b200_gemm_set_tma_store(1)
    if "--cublas" in sys.argv:
        run_cublas(name, N, K)
