"""Where a tcgen05 GEMM CTA spends its time (clock64 counters of CTA 0), config-2 shapes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from new_vit_b200 import _cabi
L = _cabi.lib()
M = int(os.environ.get("M", 526336))
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, N, K, mode in [("qkv", 1152, 384, 0), ("proj", 384, 384, 2), ("fc1", 1536, 384, 1), ("fc2", 384, 1536, 2)]:
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    b = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda").bfloat16()
    dbg = torch.zeros(96, dtype=torch.int64, device="cuda")
    for _ in range(2):
        dbg.zero_()
        _cabi.check(L.mst_debug_gemm_timing(_cabi.ptr(A), _cabi.ptr(W), M, N, K, mode, _cabi.ptr(b), _cabi.ptr(out), _cabi.ptr(out), _cabi.ptr(dbg), st))
    torch.cuda.synchronize()
    d = dbg.cpu().tolist()
    t = max(d[3], 1)
    print(f"{name:5s}: per tile (cycles) total {d[2]/t:.0f} | MMA warp: wait accumulator {d[0]/t:.0f}, wait operands {d[1]/t:.0f}, issue+rest {(d[2]-d[0]-d[1])/t:.0f}"
          f" | epilogue warp 0: wait MMA {d[4]/t:.0f}, TMEM read {d[5]/t:.0f}, math+store {d[6]/t:.0f}"
          f" (math {d[8]/t:.0f}, wait staging tile {d[9]/t:.0f}, st.shared+fence {d[10]/t:.0f}, TMA issue {d[11]/t:.0f})"
          + ("  [gemm_wt: warp 0 handles every other tile; phases are wait / math+sts / fence / TMA issue]" if K == 384 and N >= 1024 and os.environ.get("MST_GEMM_WT", "1") != "0" else ""))
    if any(d[16:]):
        tt = max(d[3] // 2, 1)
        for cta in (0, 1):
            print("   CTA", cta, "per own tile, busy:", [round(v / tt) for v in d[16 + cta * 32: 32 + cta * 32]])
            print("   CTA", cta, "per own tile, wait:", [round(v / tt) for v in d[32 + cta * 32: 48 + cta * 32]])
    del A, W, out
