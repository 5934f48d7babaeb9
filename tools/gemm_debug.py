"""GPU bring-up helper for the tcgen05 GEMM: prints error statistics and timing per shape."""
import ctypes as C
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clip_retrieval_b200 as m
from clip_retrieval_b200._lib import lib, check


def run(M, N, K, act=0, bias=True, res=True, time_it=False):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    R = torch.randn(M, N, device="cuda", generator=g).bfloat16() if res else None
    Cc = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    check(lib.b200_gemm_bf16_device(A.data_ptr(), W.data_ptr(), b.data_ptr() if bias else None,
                                    R.data_ptr() if res else None, Cc.data_ptr(), M, N, K, act, 0, st), "gemm")
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    if bias:
        ref = ref + b
    if act == 1:
        ref = ref * torch.sigmoid(1.702 * ref)
    elif act == 2:
        ref = torch.nn.functional.gelu(ref)
    if res:
        ref = ref + R.float()
    err = (Cc.float() - ref).abs()
    tol = 0.01 * ref.abs() + 0.02
    bad = (err > tol) | torch.isnan(Cc.float())
    print("M=%d N=%d K=%d act=%d: max_err=%.4g mean_err=%.4g bad=%d/%d nan=%d" % (
        M, N, K, act, err.nan_to_num(1e9).max().item(), err.nan_to_num(0).mean().item(), int(bad.sum()), M * N,
        int(torch.isnan(Cc.float()).sum())))
    if bad.any():
        idx = bad.nonzero()
        print("  first bad:", idx[:8].tolist())
        rows_bad = bad.any(dim=1).nonzero().flatten()
        cols_bad = bad.any(dim=0).nonzero().flatten()
        print("  bad rows: n=%d min=%d max=%d  bad cols: n=%d min=%d max=%d" % (
            rows_bad.numel(), rows_bad.min(), rows_bad.max(), cols_bad.numel(), cols_bad.min(), cols_bad.max()))
        print("  got[0,:8]", Cc[0, :8].float().tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())
    if time_it:
        for _ in range(3):
            lib.b200_gemm_bf16_device(A.data_ptr(), W.data_ptr(), None, None, Cc.data_ptr(), M, N, K, 0, 0, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 10
        for _ in range(reps):
            lib.b200_gemm_bf16_device(A.data_ptr(), W.data_ptr(), None, None, Cc.data_ptr(), M, N, K, 0, 0, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("  b200 gemm: %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9))
        Cf = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            torch.matmul(A, W.t(), out=Cf)
        e0.record()
        for _ in range(reps):
            torch.matmul(A, W.t(), out=Cf)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("  cuBLAS   : %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9))
    return int(bad.sum()) == 0


if __name__ == "__main__":
    ok = True
    if len(sys.argv) > 1:
        lib.b200_gemm_set_pair_mode(int(sys.argv[1]))
        print("pair mode", sys.argv[1])
    ok &= run(128, 256, 64, bias=False, res=False)
    ok &= run(128, 256, 256, bias=False, res=False)
    ok &= run(256, 512, 128, bias=True, res=False)
    ok &= run(257, 768, 768, act=1)
    ok &= run(1000, 3072, 1024, act=2)
    ok &= run(300, 1280, 1280, act=0)      # BN=128 path (1280 % 256 != 0)
    ok &= run(100, 512, 640, act=0)        # small problem -> BN=128
    ok &= run(16448, 4096, 1024, act=1, time_it=True)
    ok &= run(16448, 1024, 4096, act=0, time_it=True)
    ok &= run(65792, 3072, 1024, act=0, res=False, time_it=True)
    ok &= run(263168, 1024, 1024, act=0, res=True, time_it=True)
    ok &= run(4112, 1024, 4096, act=0, time_it=True)
    ok &= run(20000, 768, 3072, act=2, time_it=True)
    print("GEMM_DEBUG", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)
