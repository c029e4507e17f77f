"""Kernel time of the tcgen05 attention kernel alone (five launches back to back), config-2 shape by default."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from new_vit_b200 import _cabi
L = _cabi.lib()
BD, heads, N = int(os.environ.get("BD", 2048)), int(os.environ.get("HEADS", 6)), int(os.environ.get("NTOK", 257))
E = heads * 64
qkv = (torch.randn(BD * N, 3 * E, device="cuda") * 0.5).bfloat16()
out = torch.empty(BD * N, E, device="cuda").bfloat16()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    _cabi.check(L.mst_kernel_attention_bf16(_cabi.ptr(qkv), _cabi.ptr(out), BD, N, heads, st))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    _cabi.check(L.mst_kernel_attention_bf16(_cabi.ptr(qkv), _cabi.ptr(out), BD, N, heads, st))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
flops = 4.0 * BD * heads * N * N * 64
print(f"attention BD={BD} heads={heads} N={N}: {ms:.3f} ms per launch, {flops / ms / 1e9:.0f} TFLOP/s")
