"""Micro-benchmark of the tcgen05 GEMM through the C ABI (mst_kernel_gemm_bf16), config-2 shapes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from new_vit_b200 import _cabi
L = _cabi.lib()
M = int(os.environ.get("M", 526336))
shapes = [("qkv", 1152, 384, 0), ("proj", 384, 384, 2), ("fc1", 1536, 384, 1), ("fc2", 384, 1536, 2)]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, N, K, mode in shapes:
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    b = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda").bfloat16()
    res = out
    for _ in range(3):
        _cabi.check(L.mst_kernel_gemm_bf16(_cabi.ptr(A), _cabi.ptr(W), M, N, K, mode, _cabi.ptr(b), _cabi.ptr(res), _cabi.ptr(out), st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        _cabi.check(L.mst_kernel_gemm_bf16(_cabi.ptr(A), _cabi.ptr(W), M, N, K, mode, _cabi.ptr(b), _cabi.ptr(res), _cabi.ptr(out), st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:5s} M={M} N={N} K={K} mode={mode}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.0f} TFLOP/s  BN={os.environ.get('MST_GEMM_BN','192')}")
    if mode in (0, 1):   # the same GEMM with the LayerNorm folded into its epilogue (EPI_LN_*)
        stat = torch.empty(M, device="cuda")
        _cabi.check(L.mst_kernel_row_stats_bf16(_cabi.ptr(A), _cabi.ptr(stat), M, K, 1e-6, st))
        for _ in range(3):
            _cabi.check(L.mst_kernel_gemm_bf16_ln(_cabi.ptr(A), _cabi.ptr(W), M, N, K, mode, _cabi.ptr(b), _cabi.ptr(stat), _cabi.ptr(out), st))
        e0.record()
        for _ in range(n):
            _cabi.check(L.mst_kernel_gemm_bf16_ln(_cabi.ptr(A), _cabi.ptr(W), M, N, K, mode, _cabi.ptr(b), _cabi.ptr(stat), _cabi.ptr(out), st))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{name:5s} +LN fold: {ms:.3f} ms  {2*M*N*K/ms/1e9:.0f} TFLOP/s")
        e0.record()
        for _ in range(n):
            _cabi.check(L.mst_kernel_row_stats_bf16(_cabi.ptr(A), _cabi.ptr(stat), M, K, 1e-6, st))
        e1.record(); torch.cuda.synchronize()
        print(f"      row stats: {e0.elapsed_time(e1) / n:.3f} ms")
    del A, W, out
