"""Kernel-level parity of the encoder backward pieces (csrc/train_enc.cu + the fp32-output mode of the tcgen05 GEMM), each through
the C ABI against torch autograd in fp32 on the same (bf16-rounded) inputs."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from new_vit_b200 import _cabi
    return _cabi, _cabi.lib()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _relerr(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


@pytest.mark.parametrize("BD,N,heads", [(2, 257, 6), (3, 65, 6), (1, 325, 12), (5, 17, 6)])
def test_attention_backward(BD, N, heads):
    """attention.py:56-69 backward: dq (w.r.t. the un-scaled q), dk, dv from q' = q/8, k, v, o and dO; P recomputed in registers."""
    cabi, L = _lib()
    E = heads * 64
    g = torch.Generator(device="cuda").manual_seed(N * 7 + BD)
    qkv = torch.randn(BD * N, 3 * E, device="cuda", generator=g)
    qkv[:, :E] *= 0.6
    qkv = qkv.bfloat16()
    dO = (torch.randn(BD * N, E, device="cuda", generator=g) * 0.5).bfloat16()
    # fp32 autograd reference on the bf16 values; q_raw = 8 q' so that q' = q_raw / 8 as attention.py:60 scales it
    q8, k, v = [t.clone().requires_grad_(True) for t in qkv.float().reshape(BD, N, 3, heads, 64).permute(2, 0, 3, 1, 4)]
    q_raw = (q8.detach() * 8.0).requires_grad_(True)
    p = ((q_raw * 0.125) @ k.transpose(-2, -1)).softmax(-1)
    o = (p @ v).transpose(1, 2).reshape(BD * N, E)
    o.backward(dO.float())
    want = torch.stack([q_raw.grad, k.grad, v.grad]).permute(1, 3, 0, 2, 4).reshape(BD * N, 3 * E)   # [BD, N, 3, heads, 64]
    got = torch.full((BD * N, 3 * E), float("nan"), device="cuda", dtype=torch.bfloat16)
    cabi.check(L.mst_kernel_attention_bwd_bf16(cabi.ptr(qkv), cabi.ptr(o.detach().bfloat16().contiguous()), cabi.ptr(dO), cabi.ptr(got),
                                               BD, N, heads, _stream()))
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all()
    for i, name in enumerate(("dq", "dk", "dv")):
        a, b = got.float()[:, i * E:(i + 1) * E], want[:, i * E:(i + 1) * E]
        # P and dS pass through bf16 before the second product: ~2^-8 relative per term
        assert _relerr(a, b) <= 2e-2, (name, _relerr(a, b))
        torch.testing.assert_close(a, b, rtol=5e-2, atol=2e-2 * float(b.abs().max()))


def test_forward_keeps_the_row_log_sum_exp_for_the_backward():
    """N = 257: the tcgen05 forward kernel writes lse = log2 sum_j exp(s_ij) per (slice, head, token); the backward pass fed with it
    gives the gradients of the recomputing path (same kernel, phase 1 skipped)."""
    cabi, L = _lib()
    BD, N, heads = 3, 257, 6
    E = heads * 64
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = torch.randn(BD * N, 3 * E, device="cuda", generator=g)
    qkv[:, :E] *= 0.6
    qkv[5, :E] *= 6.0                      # one peaky query row
    qkv = qkv.bfloat16()
    dO = (torch.randn(BD * N, E, device="cuda", generator=g) * 0.5).bfloat16()
    o = torch.empty(BD * N, E, device="cuda", dtype=torch.bfloat16)
    lse = torch.full((BD * heads, N), float("nan"), device="cuda")
    cabi.check(L.mst_kernel_attention_lse_bf16(cabi.ptr(qkv), cabi.ptr(o), cabi.ptr(lse), BD, heads, _stream()))
    q, k, _ = qkv.float().reshape(BD, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    want = torch.logsumexp(q @ k.transpose(-2, -1), dim=-1).reshape(BD * heads, N) * 1.4426950408889634
    torch.testing.assert_close(lse, want, rtol=0, atol=2e-3)
    o2 = torch.empty_like(o)
    cabi.check(L.mst_kernel_attention_bf16(cabi.ptr(qkv), cabi.ptr(o2), BD, N, heads, _stream()))
    assert torch.equal(o, o2)              # the output does not depend on whether lse is kept
    a = torch.full((BD * N, 3 * E), float("nan"), device="cuda", dtype=torch.bfloat16)
    b = torch.full_like(a, float("nan"))
    cabi.check(L.mst_kernel_attention_bwd_bf16(cabi.ptr(qkv), cabi.ptr(o), cabi.ptr(dO), cabi.ptr(a), BD, N, heads, _stream()))
    cabi.check(L.mst_kernel_attention_bwd_lse_bf16(cabi.ptr(qkv), cabi.ptr(o), cabi.ptr(dO), cabi.ptr(lse), cabi.ptr(b), BD, N, heads, _stream()))
    torch.cuda.synchronize()
    assert torch.isfinite(b.float()).all()
    assert _relerr(b.float(), a.float()) <= 4e-3


@pytest.mark.parametrize("rows,E", [(1000, 384), (7, 384), (4099, 768)])
@pytest.mark.parametrize("with_res", [False, True])
def test_layernorm_backward(rows, E, with_res):
    cabi, L = _lib()
    g = torch.Generator(device="cuda").manual_seed(rows + E)
    x = (torch.randn(rows, E, device="cuda", generator=g) * 2 + 0.5).bfloat16()
    dy = torch.randn(rows, E, device="cuda", generator=g).bfloat16()
    dres = torch.randn(rows, E, device="cuda", generator=g).bfloat16() if with_res else None
    gamma = (1 + 0.2 * torch.randn(E, device="cuda", generator=g)).requires_grad_(True)
    beta = torch.zeros(E, device="cuda", requires_grad=True)
    xr = x.float().requires_grad_(True)
    F.layer_norm(xr, (E,), gamma, beta, 1e-6).backward(dy.float())
    want_dx = xr.grad + (dres.float() if with_res else 0)
    dx = torch.empty_like(x)
    dg, db = torch.empty(E, device="cuda"), torch.empty(E, device="cuda")
    cabi.check(L.mst_kernel_ln_bwd_bf16(cabi.ptr(x), cabi.ptr(dy), cabi.ptr(dres), cabi.ptr(gamma.detach()), cabi.ptr(dx), cabi.ptr(dg),
                                        cabi.ptr(db), rows, E, 1e-6, _stream()))
    torch.cuda.synchronize()
    torch.testing.assert_close(dx.float(), want_dx, rtol=8e-3, atol=8e-3)
    torch.testing.assert_close(dg, gamma.grad, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(db, beta.grad, rtol=1e-4, atol=1e-3)


def test_gelu_forward_and_backward():
    cabi, L = _lib()
    n = 8 * 12345
    u = (torch.randn(n, device="cuda") * 2.5).bfloat16()
    dy = torch.randn(n, device="cuda").bfloat16()
    y, du = torch.empty_like(u), torch.empty_like(u)
    cabi.check(L.mst_kernel_gelu_bf16(cabi.ptr(u), cabi.ptr(y), cabi.ptr(dy), cabi.ptr(du), n, _stream()))
    torch.cuda.synchronize()
    ur = u.float().requires_grad_(True)
    F.gelu(ur).backward(dy.float())
    torch.testing.assert_close(y.float(), F.gelu(u.float()), rtol=8e-3, atol=1e-3)
    torch.testing.assert_close(du.float(), ur.grad, rtol=8e-3, atol=4e-3)


@pytest.mark.parametrize("M,C", [(257, 384), (1000, 1536), (64, 64), (8224, 1152)])
def test_transpose_with_column_sums(M, C):
    cabi, L = _lib()
    Mpad = (M + 63) // 64 * 64
    x = torch.randn(M, C, device="cuda").bfloat16()
    out = torch.full((C, Mpad), float("nan"), device="cuda", dtype=torch.bfloat16)
    cs = torch.zeros(C, device="cuda")
    cabi.check(L.mst_kernel_transpose_bf16(cabi.ptr(x), cabi.ptr(out), cabi.ptr(cs), M, C, Mpad, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(out[:, :M], x.t()) and bool((out[:, M:] == 0).all())
    torch.testing.assert_close(cs, x.float().sum(0), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("Nout,Kin,M", [(1536, 384, 8256), (384, 1536, 4160), (1152, 384, 1024), (384, 384, 65792)])
def test_weight_gradient_gemm_fp32_out(Nout, Kin, M):
    """dW [Nout, Kin] = dY^T X as the tcgen05 GEMM sees it: A = dY^T [Nout, M], weight = X^T [Kin, M], contraction over the tokens."""
    cabi, L = _lib()
    g = torch.Generator(device="cuda").manual_seed(Nout + Kin + M)
    dY = (torch.randn(M, Nout, device="cuda", generator=g) * 0.1).bfloat16()
    X = torch.randn(M, Kin, device="cuda", generator=g).bfloat16()
    dYt, Xt = dY.t().contiguous(), X.t().contiguous()
    out = torch.full((Nout, Kin), float("nan"), device="cuda")
    cabi.check(L.mst_kernel_gemm_bf16_f32out(cabi.ptr(dYt), cabi.ptr(Xt), Nout, Kin, M, cabi.ptr(out), _stream()))
    torch.cuda.synchronize()
    want = dY.float().t() @ X.float()
    torch.testing.assert_close(out, want, rtol=2e-3, atol=2e-3 * float(want.abs().max()))


@pytest.mark.parametrize("Nout,Kin,M", [(1536, 384, 8256), (384, 1536, 4160), (1152, 384, 1000), (384, 384, 65792), (128, 192, 37),
                                        (2304, 768, 5200)])
@pytest.mark.parametrize("bias", [True, False])
def test_weight_gradient_from_row_major_activations(Nout, Kin, M, bias):
    """dW = dY^T X and db = colsum(dY) read from the row-major activations (MN-major tcgen05 operands; ragged token tails are
    zero-filled by TMA).  Outputs are overwritten, not accumulated."""
    cabi, L = _lib()
    g = torch.Generator(device="cuda").manual_seed(3 * Nout + Kin + M)
    dY = (torch.randn(M, Nout, device="cuda", generator=g) * 0.1).bfloat16()
    X = torch.randn(M, Kin, device="cuda", generator=g).bfloat16()
    dW = torch.full((Nout, Kin), float("nan"), device="cuda")
    db = torch.full((Nout,), float("nan"), device="cuda") if bias else None
    cabi.check(L.mst_kernel_wgrad_bf16(cabi.ptr(dY), cabi.ptr(X), M, Nout, Kin, cabi.ptr(dW), cabi.ptr(db) if bias else None, _stream()))
    torch.cuda.synchronize()
    want = dY.float().t() @ X.float()
    torch.testing.assert_close(dW, want, rtol=2e-3, atol=2e-3 * float(want.abs().max()))
    if bias:
        wb = dY.float().sum(0)
        torch.testing.assert_close(db, wb, rtol=1e-4, atol=1e-4 * float(wb.abs().max()) + 1e-4)

