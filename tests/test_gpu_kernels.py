"""Kernel-level parity (GPU): each CUDA kernel, called through the C ABI, against a plain PyTorch fp32
reference of the same op on the same inputs."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from new_vit_b200 import _cabi
    return _cabi, _cabi.lib()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _gelu(x):
    return torch.nn.functional.gelu(x)


@pytest.mark.parametrize("M,N,K", [(128, 192, 64), (1000, 384, 384), (647, 1152, 384), (300, 1536, 384),
                                   (520, 384, 1536), (64, 384, 256), (4112, 1152, 384), (40000, 384, 384),
                                   (257, 256, 128), (20000, 384, 1536),
                                   # gemm_wt (weights in TMEM; K = 384, N % 256 == 0): one token, a ragged second tile, and enough
                                   # tiles per CTA pair for the operand ring and both accumulator stages to wrap many times
                                   (1, 1536, 384), (129, 1024, 384), (70000, 1536, 384)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gemm_bf16_tcgen05(M, N, K, mode):
    cabi, L = _lib()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K + mode)
    A = (torch.randn(M, K, device="cuda", generator=g) * 1.0).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    out = torch.full((M, N), float("nan"), device="cuda").bfloat16()
    cabi.check(L.mst_kernel_gemm_bf16(cabi.ptr(A), cabi.ptr(W), M, N, K, mode, cabi.ptr(bias), cabi.ptr(res),
                                      cabi.ptr(out), _stream()))
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias
    if mode == 1:
        ref = _gelu(ref)
    if mode == 2:
        ref = ref + res.float()
    # fp32 accumulate, one bf16 rounding at the end: 2^-8 relative
    torch.testing.assert_close(out.float(), ref, rtol=8e-3, atol=8e-3)


def test_gemm_bf16_inplace_residual():
    """x += A W^T + b in place: the update is a TMA reduce-add, i.e. the UPDATE is rounded to bf16 before the memory system adds it
    to the bf16 residual (two roundings).  The bound therefore carries half an ulp of the larger operand next to the usual output
    rounding: where x and the update nearly cancel, the error is set by their magnitude, not by the result's.  (Unseeded and with a
    flat 8e-3 this test failed about one run in ten on exactly such an element.)"""
    cabi, L = _lib()
    M, N, K = 777, 384, 384
    g = torch.Generator(device="cuda").manual_seed(777)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    x0 = x.float()
    delta = A.float() @ W.float().t() + bias
    ref = x0 + delta
    cabi.check(L.mst_kernel_gemm_bf16(cabi.ptr(A), cabi.ptr(W), M, N, K, 2, cabi.ptr(bias), cabi.ptr(x), cabi.ptr(x), _stream()))
    torch.cuda.synchronize()
    bound = 8e-3 * (1.0 + ref.abs()) + 2.0 ** -8 * torch.maximum(x0.abs(), delta.abs())
    err = (x.float() - ref).abs()
    assert bool((err <= bound).all()), f"max excess {(err - bound).max().item():.3e}"


@pytest.mark.parametrize("M,N,K", [(130, 384, 384), (257, 1152, 384), (100, 384, 1536), (65, 384, 256)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gemm_f32(M, N, K, mode):
    cabi, L = _lib()
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") * K ** -0.5
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda")
    cabi.check(L.mst_kernel_gemm_f32(cabi.ptr(A), cabi.ptr(W), M, N, K, mode, cabi.ptr(bias), cabi.ptr(res), cabi.ptr(out), _stream()))
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().t() + bias.double())
    if mode == 1:
        ref = _gelu(ref)
    if mode == 2:
        ref = ref + res.double()
    torch.testing.assert_close(out.double(), ref, rtol=1e-5, atol=1e-5)


def _attn_ref(qkv, BD, N, heads):
    E = heads * 64
    q, k, v = qkv.float().reshape(BD, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    p = (q @ k.transpose(-2, -1)).softmax(-1)  # q is pre-scaled
    return (p @ v).transpose(1, 2).reshape(BD * N, E)


# N == 257: attention_tc257_kernel; 17 <= N <= 352: attention_tcg_kernel (one or two S chunks, one to four query tiles,
# ragged last tile / key padding); N == 16 and N == 400: the warp-MMA kernel.  (40, 325, 12) gives every SM more than one item.
@pytest.mark.parametrize("BD,N,heads", [(3, 257, 6), (2, 65, 6), (2, 325, 12), (1, 16, 6), (2, 33, 6), (1, 17, 6), (3, 109, 6),
                                        (2, 261, 6), (40, 325, 12), (2, 352, 6), (1, 400, 6), (70, 129, 6)])
def test_attention_bf16(BD, N, heads):
    cabi, L = _lib()
    E = heads * 64
    g = torch.Generator(device="cuda").manual_seed(N)
    qkv = torch.randn(BD * N, 3 * E, device="cuda", generator=g)
    qkv[:, :E] *= 0.5  # sharper than uniform, like a scaled q
    qkv = qkv.bfloat16()
    out = torch.full((BD * N, E), float("nan"), device="cuda").bfloat16()
    cabi.check(L.mst_kernel_attention_bf16(cabi.ptr(qkv), cabi.ptr(out), BD, N, heads, _stream()))
    torch.cuda.synchronize()
    ref = _attn_ref(qkv, BD, N, heads)
    # P is rounded to bf16 before P.V and the output to bf16: ~2^-8 relative on O(1) values
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("BD,heads", [(1, 6), (25, 6), (130, 6), (40, 12)])
def test_attention_tcgen05_257(BD, heads):
    """N == 257 routes to the tcgen05 kernel; many items per CTA exercise the stage ring and barrier phases."""
    cabi, L = _lib()
    N, E = 257, heads * 64
    g = torch.Generator(device="cuda").manual_seed(BD)
    qkv = torch.randn(BD * N, 3 * E, device="cuda", generator=g)
    qkv[:, :E] *= 0.6
    qkv = qkv.bfloat16()
    out = torch.full((BD * N, E), float("nan"), device="cuda").bfloat16()
    out2 = torch.full((BD * N, E), float("nan"), device="cuda").bfloat16()
    cabi.check(L.mst_kernel_attention_bf16(cabi.ptr(qkv), cabi.ptr(out), BD, N, heads, _stream()))
    cabi.check(L.mst_kernel_attention_bf16_warp_mma(cabi.ptr(qkv), cabi.ptr(out2), BD, N, heads, _stream()))
    torch.cuda.synchronize()
    ref = _attn_ref(qkv, BD, N, heads)
    torch.testing.assert_close(out2.float(), ref, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("BD,N,heads", [(2, 257, 6), (1, 65, 6), (1, 325, 12)])
def test_attention_f32(BD, N, heads):
    cabi, L = _lib()
    E = heads * 64
    qkv = torch.randn(BD * N, 3 * E, device="cuda")
    out = torch.full((BD * N, E), float("nan"), device="cuda")
    cabi.check(L.mst_kernel_attention_f32(cabi.ptr(qkv), cabi.ptr(out), BD, N, heads, _stream()))
    torch.cuda.synchronize()
    torch.testing.assert_close(out, _attn_ref(qkv, BD, N, heads), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("rows,E", [(1000, 384), (7, 384), (513, 768)])
def test_layernorm_bf16(rows, E):
    cabi, L = _lib()
    x = (torch.randn(rows, E, device="cuda") * 3 + 1).bfloat16()
    g = torch.randn(E, device="cuda")
    b = torch.randn(E, device="cuda")
    y = torch.empty_like(x)
    cabi.check(L.mst_kernel_layernorm_bf16(cabi.ptr(x), cabi.ptr(y), cabi.ptr(g), cabi.ptr(b), rows, E, 1e-6, _stream()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float(), (E,), g, b, 1e-6)
    torch.testing.assert_close(y.float(), ref, rtol=8e-3, atol=8e-3)


def test_saliency_upsample_matches_trilinear():
    cabi, L = _lib()
    B, D, heads, sheads, g = 2, 5, 6, 12, 16
    plane = torch.rand(B * D, heads, g * g + 1, device="cuda")
    plane = plane / plane.sum(-1, keepdim=True)
    slc = torch.rand(B, sheads, D + 1, device="cuda")
    slc = slc / slc.sum(-1, keepdim=True)
    H = W = 224
    maps = torch.empty(B * D, heads, g * g, device="cuda")
    pl = torch.empty_like(maps)
    sl = torch.empty(B * D, device="cuda")
    coarse = torch.empty(B, 1, D, g, g, device="cuda")
    full = torch.empty(B, 1, D, H, W, device="cuda")
    cabi.check(L.mst_saliency(None, cabi.ptr(plane), cabi.ptr(slc), B, D, heads, sheads, 1, g, g, H, W, 0, cabi.ptr(maps), cabi.ptr(pl),
                              cabi.ptr(sl), cabi.ptr(coarse), cabi.ptr(full), _stream()))
    torch.cuda.synchronize()
    from oracle import mst_oracle as O
    rp, rs = plane.cpu(), slc.cpu()
    torch.testing.assert_close(maps.cpu(), O.get_attention_maps(rp, rs), rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(pl.cpu(), O.get_plane_attention(rp), rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(sl.cpu(), O.get_slice_attention(rs).reshape(-1), rtol=1e-5, atol=1e-9)
    rc, rf, _ = O.saliency(rp, rs, B, D, H, W)
    torch.testing.assert_close(coarse.cpu(), rc, rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(full.cpu(), rf, rtol=1e-5, atol=float(rf.max()) * 1e-6)
    # indexing is bit-exact: same argmax voxel
    assert torch.equal(full.cpu().reshape(B, -1).argmax(-1), rf.reshape(B, -1).argmax(-1))


@pytest.mark.parametrize("n,items", [(1, 1), (2, 3), (1000, 2), (224 * 224 * 8 + 3, 2)])
def test_quantile_matches_numpy_bit_exact(n, items):
    """np.quantile (numpy 'linear' method; scripts/main_predict.py:243-245,296): radix select + numpy's own lerp arithmetic.
    Ties, negatives, zeros of both signs and the q = 0 / 1 ends included."""
    import numpy as np
    from new_vit_b200.model import quantile
    g = torch.Generator().manual_seed(n)
    x = torch.randn(items, n, generator=g)
    x[:, ::7] = x[:, :1].clone()    # ties
    if n > 4:
        x[0, 1], x[0, 2] = 0.0, -0.0
    qs = [0.0, 0.25, 0.5, 0.995, 0.999, 1.0]
    got = quantile(x.cuda(), qs).cpu().numpy()
    want = np.stack([np.quantile(r, qs) for r in x.numpy()])
    assert got.shape == want.shape and np.array_equal(got, want), (got, want)


def test_rollout_matches_matmul_chain():
    """get_attention_cls (dino.py:204-212): R = maps[-1]; for attn in reversed(maps[:-1]): R = attn @ R."""
    cabi, L = _lib()
    depth, nmat, N = 5, 7, 67   # N not a multiple of the 64-wide tile
    maps = torch.rand(depth, nmat, N, N, device="cuda")
    maps = maps / maps.sum(-1, keepdim=True)
    out, scratch = torch.empty(nmat, N, N, device="cuda"), torch.empty(nmat, N, N, device="cuda")
    for d in (depth, 2, 1):
        cabi.check(L.mst_rollout(cabi.ptr(maps), d, nmat, N, cabi.ptr(out), cabi.ptr(scratch), _stream()))
        torch.cuda.synchronize()
        ref = maps[d - 1].double().cpu()
        for l in range(d - 2, -1, -1):
            ref = maps[l].double().cpu() @ ref
        torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("M,N,K", [(1000, 1152, 384), (4112, 1536, 384), (333, 2304, 768), (300, 3072, 768), (33, 1536, 384),
                                   (50001, 1536, 384)])
@pytest.mark.parametrize("gelu", [0, 1])
def test_gemm_bf16_layernorm_folded(M, N, K, gelu):
    """norm1 -> qkv and norm2 -> fc1 with the LayerNorm folded into the GEMM (block.py:112-113): raw rows through the tensor
    cores against gamma-scaled, row-centred weights, per-row rstd applied in the epilogue.  Reference: LN in fp32, then the Linear."""
    cabi, L = _lib()
    g = torch.Generator(device="cuda").manual_seed(M + N + K + gelu)
    x = (torch.randn(M, K, device="cuda", generator=g) * 1.7 + 0.9).bfloat16()   # non-zero mean: the -mean*csum term matters
    x[:, 5] += 20.0                                                               # an outlier channel, as real ViTs have
    W0 = torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5)
    b0 = torch.randn(N, device="cuda", generator=g)
    gamma = 1.0 + 0.2 * torch.randn(K, device="cuda", generator=g)
    beta = 0.1 * torch.randn(K, device="cuda", generator=g)
    Wg = W0 * gamma
    Wf = (Wg - Wg.mean(-1, keepdim=True)).bfloat16()     # centred rows: the row mean of x cancels inside the MMA
    bias = (b0 + W0 @ beta).contiguous()
    stat = torch.empty(M, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda").bfloat16()
    cabi.check(L.mst_kernel_row_stats_bf16(cabi.ptr(x), cabi.ptr(stat), M, K, 1e-6, _stream()))
    cabi.check(L.mst_kernel_gemm_bf16_ln(cabi.ptr(x), cabi.ptr(Wf), M, N, K, gelu, cabi.ptr(bias), cabi.ptr(stat), cabi.ptr(out),
                                         _stream()))
    torch.cuda.synchronize()
    xf = x.float()
    var, mean = torch.var_mean(xf, dim=-1, unbiased=False, keepdim=True)
    torch.testing.assert_close(stat[:, None], torch.rsqrt(var + 1e-6), rtol=1e-5, atol=0)
    ref = torch.nn.functional.layer_norm(xf, (K,), gamma, beta, 1e-6) @ W0.t() + b0
    if gelu:
        ref = _gelu(ref)
    torch.testing.assert_close(out.float(), ref, rtol=1.2e-2, atol=1.2e-2)


@pytest.mark.parametrize("M", [777, 4112, 70000])
@pytest.mark.parametrize("case", ["plain", "large_mean", "outlier_channel"])
def test_gemm_fc2_fused_row_statistics(M, case):
    """fc2 between two blocks (block.py:113 -> the next block's norm1, :112): x += A W^T + b in place and rstd of the UPDATED rows
    out of the same epilogue (one-pass sum / sum of squares of the bf16 values it writes).  Against two-pass fp32 statistics of
    the rows the kernel wrote, on rows with |mean| = 20 std and with a 100x outlier channel (massive activations)."""
    cabi, L = _lib()
    N, K = 384, 1536
    g = torch.Generator(device="cuda").manual_seed(M + len(case))
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g)
    if case == "large_mean":
        x = x + 20.0 * (1.0 + torch.rand(M, 1, device="cuda", generator=g))     # row mean 20..40 x the row std
    if case == "outlier_channel":
        x[:, 7] += 100.0
        x[:, 300] -= 60.0
    x = x.bfloat16()
    ref = x.float() + A.float() @ W.float().t() + bias
    stat = torch.full((M,), float("nan"), device="cuda")
    cabi.check(L.mst_kernel_gemm_bf16_res_stats(cabi.ptr(A), cabi.ptr(W), M, N, K, cabi.ptr(bias), cabi.ptr(x), cabi.ptr(stat),
                                                1e-6, _stream()))
    torch.cuda.synchronize()
    bound = 8e-3 * (1.0 + ref.abs())
    assert bool(((x.float() - ref).abs() <= bound).all())
    var, _ = torch.var_mean(x.float(), dim=-1, unbiased=False)       # two-pass fp32 on what was written
    want = torch.rsqrt(var + 1e-6)
    # half a bf16 ulp (2^-9 = 2e-3) is what the consuming GEMM's output rounding costs anyway; hold the statistics to 1e-3
    torch.testing.assert_close(stat, want, rtol=1e-3, atol=0)


@pytest.mark.parametrize("N,K", [(1152, 384), (1536, 384), (2304, 768)])
def test_ln_fold_packing_keeps_rows_centred_after_bf16_rounding(N, K):
    """norm1 -> qkv / norm2 -> fc1 folding (api.cu pack_linear_ln_kernel): the bf16 weight rows must still sum to ~0, otherwise
    mean(x) * rstd * sum_k Wc[n,k] leaks into every output.  Rows with mean = 20 std through the packed weights + folded GEMM
    against LayerNorm-then-Linear in fp32."""
    cabi, L = _lib()
    g = torch.Generator(device="cuda").manual_seed(N + K)
    W0 = torch.randn(N, K, device="cuda", generator=g).clamp(-2, 2) * 0.02          # trunc_normal(std 0.02)
    b0 = 0.1 * torch.randn(N, device="cuda", generator=g)
    gamma = 1.0 + 0.3 * torch.randn(K, device="cuda", generator=g)
    beta = 0.1 * torch.randn(K, device="cuda", generator=g)
    Wd = torch.empty(N, K, device="cuda", dtype=torch.bfloat16)
    bd = torch.empty(N, device="cuda")
    cabi.check(L.mst_kernel_pack_linear_ln(cabi.ptr(W0), cabi.ptr(b0), cabi.ptr(gamma), cabi.ptr(beta), N, K, cabi.ptr(Wd), cabi.ptr(bd),
                                           _stream()))
    torch.cuda.synchronize()
    Wg = W0 * gamma
    Wc = Wg - Wg.mean(-1, keepdim=True)
    ulp = 2.0 ** (torch.floor(torch.log2(torch.maximum(Wc.abs(), Wd.float().abs()).clamp_min(1e-30))) - 7)
    # a moved element sits one bf16 step from its rounding: at most 1.5 of its own steps from the exact value (half a step
    # of rounding plus the move), and only a few elements per row are moved
    # (+ 1e-8: the kernel's own fp32 evaluation of W * gamma - mean differs from this one by a few 1e-9, which is many ulps of an
    #  element that happens to be ~1e-7)
    assert bool(((Wd.float() - Wc).abs() <= 1.51 * ulp + 1e-8).all())
    assert ((Wd.float() - Wc).abs() > 0.51 * ulp + 1e-8).float().mean().item() < 0.08
    rowsum = Wd.double().sum(-1).abs()
    plain = Wc.bfloat16().double().sum(-1).abs()
    assert rowsum.max().item() <= 2e-6, rowsum.max().item()                          # independent roundings leave ~7e-4 (K = 384)
    assert plain.max().item() > 50 * rowsum.max().item()
    torch.testing.assert_close(bd, b0 + W0 @ beta, rtol=1e-5, atol=1e-6)
    M = 2000
    x = (torch.randn(M, K, device="cuda", generator=g) + 20.0).bfloat16()            # row mean = 20 std
    stat = torch.empty(M, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    cabi.check(L.mst_kernel_row_stats_bf16(cabi.ptr(x), cabi.ptr(stat), M, K, 1e-6, _stream()))
    cabi.check(L.mst_kernel_gemm_bf16_ln(cabi.ptr(x), cabi.ptr(Wd), M, N, K, 0, cabi.ptr(bd), cabi.ptr(stat), cabi.ptr(out), _stream()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float(), (K,), gamma, beta, 1e-6) @ W0.t() + b0
    # the error budget is the bf16 weight rounding (|LN(x)| ~ 1 over K terms of 2^-9 relative error: ~3e-3 at 4.5 sigma) + the
    # output rounding (4e-3 at |y| ~ 2): 8e-3, the tolerance of the plain GEMM tests.  With independently rounded weight rows the
    # same rows miss it by 5x (mean(x) * rstd * sum_k Wc[n,k] ~ 20 * 1.6e-3).
    torch.testing.assert_close(out.float(), ref, rtol=8e-3, atol=8e-3)
