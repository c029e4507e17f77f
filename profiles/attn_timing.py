"""Phase timing of the tcgen05 attention kernel (one softmax warp of CTA 0) + kernel time, config-2 shape."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from new_vit_b200 import _cabi
L = _cabi.lib()
BD, heads, N = int(os.environ.get("BD", 2048)), 6, 257
E = heads * 64
qkv = (torch.randn(BD * N, 3 * E, device="cuda") * 0.5).bfloat16()
out = torch.empty(BD * N, E, device="cuda").bfloat16()
dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    dbg.zero_()
    _cabi.check(L.mst_debug_attention_timing(_cabi.ptr(qkv), _cabi.ptr(out), BD, heads, _cabi.ptr(dbg), st))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    _cabi.check(L.mst_kernel_attention_bf16(_cabi.ptr(qkv), _cabi.ptr(out), BD, N, heads, st))
e1.record(); torch.cuda.synchronize()
d = dbg.cpu().tolist()
if os.environ.get("MST_ATTN_WARPS", "16") != "8":
    # the 16-softmax-warp kernel carries no phase counters: at 80 registers per thread they spilled and cost 5-10 % in-step
    # (history: commits 105772b..871446d); phase numbers below are the 8-warp kernel's only
    print(f"kernel {e0.elapsed_time(e1)/5:.3f} ms (16 softmax warps; run with MST_ATTN_WARPS=8 for the phase breakdown of the 8-warp kernel)")
    sys.exit(0)
tiles = max(d[6] // 2, 1)  # tiles handled by this warp's team
names = ["wait S", "dot+pass1", "pair barrier", "pass 2", "wait O", "epilogue"]
print("MMA warp waits per item (cycles): kv_full", round(d[8] / (d[6] // 2)), "o_free", round(d[9] / (d[6] // 2)), "sp_done", round(d[10] / (d[6] // 2)))
print("CLS warp 2 per item it handles: wait kv_full", round(d[11] / max(d[13], 1)), "compute", round(d[12] / max(d[13], 1)), "items", d[13])
print(f"kernel {e0.elapsed_time(e1)/5:.3f} ms; per tile of this team ({tiles} tiles):", {n: round(d[i] / tiles) for i, n in enumerate(names)}, "sum", round(sum(d[:6]) / tiles))
