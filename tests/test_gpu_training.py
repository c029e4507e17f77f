"""Training step on a frozen encoder (BASELINE.json config 5, frozen-encoder construction), GPU: the CUDA forward / backward of
the slice transformer + head (csrc/train.cu) against (1) gradients the REAL reference produced (tests/golden/train_*.npz),
(2) the oracle under torch autograd on fresh inputs, and the fused AdamW kernel against torch.optim.AdamW."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _sub(t, sub):
    return t[::sub] if t.dim() == 2 and t.shape[0] >= 64 else t


@pytest.mark.parametrize("name", ["train_s_frozen_b3", "train_s_frozen_d32_b2"])
def test_training_step_matches_reference_golden(name):
    """model.train(); loss = model._step(batch) (base_model.py:148-170); loss.backward(); one AdamW step (base_model.py:103-110)."""
    from new_vit_b200 import DinoV2ClassifierSlice, synth
    from new_vit_b200.training import SLICE_PARAM_NAMES, FusedAdamW
    meta, g = load_golden(name)
    sd = synth.make_state_dict("s", 2, seed=meta["wseed"], variant="peaky", img_size=meta["H"])
    x = synth.make_volume(meta["B"], meta["D"], meta["H"], meta["W"], seed=meta["vseed"])
    mask = synth.make_padding_mask(meta["B"], meta["D"], seed=meta["vseed"]) if meta["masked"] else None
    m = DinoV2ClassifierSlice(1, 2, pretrained=False, precision="fp32", img_size=meta["H"], freeze=True,
                              optimizer_kwargs={'lr': meta["lr"], 'weight_decay': 1e-2}).cuda()
    m.load_state_dict(sd)
    m.train()
    opt = m.configure_optimizers()[0]
    assert isinstance(opt, FusedAdamW)
    opt.zero_grad()
    batch = {"source": x, "target": g["target"].cuda(), "uid": ["a"] * meta["B"]}
    if mask is not None:
        batch["src_key_padding_mask"] = mask
    loss = m.training_step(batch, 0)
    assert loss.requires_grad
    torch.testing.assert_close(loss.detach().cpu().reshape(1), g["loss"], rtol=1e-4, atol=1e-5)
    loss.backward()
    params = dict(m.named_parameters())
    for n in SLICE_PARAM_NAMES:
        got = _sub(params[n].grad.detach().cpu(), meta["sub"])
        want = g["grad." + n]
        scale = float(want.abs().max())
        torch.testing.assert_close(got, want, rtol=2e-3, atol=2e-4 * scale + 1e-9, msg=lambda s, n=n: f"grad {n}: {s}")
    assert all(p.grad is None for p in m.encoder.parameters())
    opt.step()
    for n in SLICE_PARAM_NAMES:
        got = _sub(params[n].detach().cpu(), meta["sub"])
        # one AdamW step moves every element by ~lr (|m / sqrt(v)| ~ 1 at step 1): compare the update, not the value
        before = _sub(sd[n].reshape(params[n].shape), meta["sub"])
        # (Adam normalises by sqrt(v): where the gradient is rounding noise -- the key bias, whose exact gradient is 0 because
        #  softmax rows sum to 1 -- the first step is +-lr with a random sign in the reference too; those elements are skipped)
        gw = g["grad." + n]
        live = gw.abs() > 1e-4 * gw.abs().max()
        torch.testing.assert_close((got - before)[live], (g["after." + n] - before)[live], rtol=2e-2, atol=0.05 * meta["lr"],
                                   msg=lambda s, n=n: f"step {n}: {s}")
    # eval forward after the step uses the updated weights (the handle is re-packed on demand)
    m.eval()
    with torch.no_grad():
        y = m(x, src_key_padding_mask=mask)
    m2 = DinoV2ClassifierSlice(1, 2, pretrained=False, precision="fp32", img_size=meta["H"]).cuda().eval()
    m2.load_state_dict(m.state_dict())
    with torch.no_grad():
        assert torch.equal(y, m2(x, src_key_padding_mask=mask))


@pytest.mark.parametrize("B,D,masked", [(8, 32, False), (5, 7, True), (1, 64, True)])
def test_slice_backward_matches_oracle_autograd(B, D, masked):
    """csrc/train.cu through the autograd.Function on given encoder features, against torch autograd over the oracle's slice
    transformer (oracle/mst_oracle.py) with the same parameters: logits, every parameter gradient, and d/d enc."""
    from new_vit_b200 import synth
    from new_vit_b200.training import SLICE_PARAM_NAMES, SliceHeadFunction
    from oracle import mst_oracle as O
    E, heads, C = 384, 12, 2
    sd = synth.make_state_dict("s", C, seed=61, variant="peaky")
    g = torch.Generator().manual_seed(B * 100 + D)
    enc = torch.randn(B, D, E, generator=g)
    mask = synth.make_padding_mask(B, D, seed=3) if masked else None
    if masked and B == 1:
        mask[0, D - 5:] = True
    target = torch.randint(0, C, (B,), generator=g)
    # oracle under autograd (CPU)
    ref_p = {n: sd[n].clone().requires_grad_(True) for n in SLICE_PARAM_NAMES}
    enc_ref = enc.clone().requires_grad_(True)
    full = dict(sd)
    full.update(ref_p)
    tok = torch.cat([ref_p["cls_token"].repeat(B, 1, 1), enc_ref], dim=1)
    kpm = None if mask is None else torch.cat([torch.zeros((B, 1), dtype=torch.bool), mask], dim=1)
    y, _ = O.slice_transformer(full, tok, kpm)
    logits_ref = F.linear(y[:, 0], ref_p["linear.weight"], ref_p["linear.bias"])
    F.cross_entropy(logits_ref, target).backward()
    # CUDA
    ps = [sd[n].clone().cuda().requires_grad_(True) for n in SLICE_PARAM_NAMES]
    enc_c = enc.cuda().requires_grad_(True)
    mk = None if mask is None else mask.to(torch.uint8).cuda()
    logits = SliceHeadFunction.apply(None, enc_c, mk, heads, True, *ps)
    torch.testing.assert_close(logits.detach().cpu(), logits_ref.detach(), rtol=1e-4, atol=1e-5)
    F.cross_entropy(logits, target.cuda()).backward()
    for n, p in zip(SLICE_PARAM_NAMES, ps):
        want = ref_p[n].grad
        torch.testing.assert_close(p.grad.cpu(), want, rtol=1e-3, atol=1e-5 * float(want.abs().max()) + 1e-10, msg=lambda s, n=n: f"{n}: {s}")
    want = enc_ref.grad
    torch.testing.assert_close(enc_c.grad.cpu(), want, rtol=1e-3, atol=1e-5 * float(want.abs().max()))
    if mask is not None:   # masked slices get no gradient through the attention keys/values... and none at all (they only enter there)
        assert float(enc_c.grad.cpu()[mask].abs().max()) == 0.0


def test_fused_adamw_matches_torch():
    from new_vit_b200.training import FusedAdamW
    torch.manual_seed(0)
    shapes = [(384, 384), (1152,), (2, 384), (7,), (1, 1, 384)]
    ref = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ours = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    o_ref = torch.optim.AdamW(ref, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    o_our = FusedAdamW(ours, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    for step in range(5):
        grads = [torch.randn(s, device="cuda") * (0.1 + step) for s in shapes]
        o_our.zero_grad()
        for p, q, gr in zip(ref, ours, grads):
            p.grad = gr.clone()
            q.grad.copy_(gr)
        o_ref.step()
        o_our.step()
        for p, q in zip(ref, ours):
            torch.testing.assert_close(q.detach(), p.detach(), rtol=2e-6, atol=2e-7)


def test_unsupported_training_configurations_say_so():
    """The encoder's backward pass exists for bf16 at the position table's own grid; everything else raises a clear error, and
    inference under grad mode stays the inference path."""
    from new_vit_b200 import DinoV2ClassifierSlice
    from new_vit_b200._cabi import MSTError
    m = DinoV2ClassifierSlice(1, 2, pretrained=False, precision="fp32").cuda().train()
    with pytest.raises(NotImplementedError, match="freeze=True"):
        m(torch.zeros(1, 1, 2, 224, 224))
    m = DinoV2ClassifierSlice(1, 2, pretrained=False, precision="bf16").cuda().train()
    with pytest.raises(MSTError, match="position table's own grid"):
        m(torch.zeros(1, 1, 2, 28, 28))
    m.eval()
    assert m(torch.zeros(1, 1, 2, 28, 28)).shape == (1, 2)


def _cosine(a, b):
    return float(F.cosine_similarity(a.reshape(-1).double(), b.reshape(-1).double(), dim=0))


def test_full_training_step_matches_reference_golden():
    """Every parameter trainable (the construction main_train.py:36-37,110-126 trains): bf16 CUDA forward + backward of the whole
    model against the fp32 gradients of the REAL reference (tests/golden/train_s_full_b2.npz: every 61st element of every
    gradient + its norm).  bf16 activations / activation gradients: direction (cosine) and size (norm) per tensor."""
    from new_vit_b200 import DinoV2ClassifierSlice, synth
    meta, g = load_golden("train_s_full_b2")
    sd = synth.make_state_dict("s", 2, seed=meta["wseed"], variant="peaky", img_size=meta["H"])
    x = synth.make_volume(meta["B"], meta["D"], meta["H"], meta["W"], seed=meta["vseed"])
    mask = synth.make_padding_mask(meta["B"], meta["D"], seed=meta["vseed"])
    m = DinoV2ClassifierSlice(1, 2, pretrained=False, precision="bf16").cuda()
    m.load_state_dict(sd)
    m.train()
    opt = m.configure_optimizers()[0]
    opt.zero_grad()
    batch = {"source": x, "target": g["target"].cuda(), "src_key_padding_mask": mask}
    loss = m.training_step(batch, 0)
    assert abs(float(loss.detach()) - float(g["loss"])) <= 2e-2
    loss.backward()
    params = dict(m.named_parameters())
    assert float(params["encoder.mask_token"].grad.abs().max()) == 0.0      # never read by the path: no gradient
    worst = {}
    for n in meta["trainable"]:
        got = params[n].grad.detach().cpu().reshape(-1)
        want = g["grad." + n]
        cos = _cosine(got[::meta["stride"]], want)
        ratio = float(got.norm()) / max(float(g["norm." + n]), 1e-30)
        worst[n] = (cos, ratio)
    bad = {n: v for n, v in worst.items() if not (v[0] >= 0.99 and 0.95 <= v[1] <= 1.05)
           # the key bias of every attention has an exactly-zero gradient (softmax rows sum to 1): pure rounding noise on both sides
           and not n.endswith("attn.qkv.bias") and not n.endswith("in_proj_bias")}
    assert not bad, bad
    for n in meta["trainable"]:
        if n.endswith("attn.qkv.bias"):   # ... so compare its q and v thirds only
            E = 384
            got = params[n].grad.detach().cpu()
            idx = torch.arange(0, 3 * E)[::meta["stride"]]
            keep = (idx < E) | (idx >= 2 * E)
            assert _cosine(got[idx][keep], g["grad." + n][keep]) >= 0.98, n
    # one optimizer step over all 22.5 M parameters, then a consistent eval forward
    before = params["encoder.blocks.0.3.mlp.fc1.weight"].detach().clone()
    opt.step()
    assert float((params["encoder.blocks.0.3.mlp.fc1.weight"].detach() - before).abs().max()) > 0
    m.eval()
    with torch.no_grad():
        assert torch.isfinite(m(x, src_key_padding_mask=mask)).all()


def test_vit_b_training_step_matches_oracle_autograd():
    """ViT-B/14 (config 4's encoder) with every parameter trainable: E = 768 exercises the other instantiations of the backward kernels
    (LayerNorm backward EPL = 24, weight-gradient tiles 18 x 4 / 6 x 16, 12 heads).  Reference = the oracle under torch autograd on the CPU
    (pinned to the live reference's gradients by tests/test_training_cpu.py), one small volume."""
    from test_training_cpu import oracle_full_train_step
    from new_vit_b200 import DinoV2ClassifierSlice, synth
    B, D, H = 1, 3, 112
    sd = synth.make_state_dict("b", 2, seed=5, variant="peaky", img_size=H)
    x = synth.make_volume(B, D, H, H, seed=9)
    target = torch.tensor([1])
    _, loss_ref, grads = oracle_full_train_step(sd, x, None, target)
    m = DinoV2ClassifierSlice(1, 2, pretrained=False, precision="bf16", model_size="b", img_size=H).cuda()
    m.load_state_dict(sd)
    m.train()
    opt = m.configure_optimizers()[0]
    opt.zero_grad()
    loss = m.training_step({"source": x, "target": target.cuda()}, 0)
    assert abs(float(loss.detach()) - float(loss_ref)) <= 2e-2
    loss.backward()
    bad = {}
    for n, p in m.named_parameters():
        if n == "encoder.mask_token" or n.endswith("attn.qkv.bias") or n.endswith("in_proj_bias"):
            continue     # no gradient / an exactly-zero third (see the ViT-S test)
        want = grads[n]
        got = p.grad.detach().cpu()
        if float(want.norm()) < 1e-12:
            continue
        cos, ratio = _cosine(got, want), float(got.norm()) / float(want.norm())
        if not (cos >= 0.99 and 0.95 <= ratio <= 1.05):
            bad[n] = (cos, ratio)
    assert not bad, bad


@pytest.mark.parametrize("B,D,masked", [(1, 1, False), (3, 5, True), (2, 7, False)])
def test_full_training_step_on_other_batch_shapes(B, D, masked):
    """Whole-model gradients at batch / slice counts the goldens do not cover (one slice per volume: two slice-transformer tokens; odd
    counts; padding masks), against the oracle under torch autograd (pinned to the live reference by tests/test_training_cpu.py)."""
    from test_training_cpu import oracle_full_train_step
    from new_vit_b200 import DinoV2ClassifierSlice, synth
    sd = synth.make_state_dict("s", 2, seed=21 + B, variant="peaky")
    x = synth.make_volume(B, D, 224, 224, seed=5 + D)
    mask = synth.make_padding_mask(B, D, seed=B) if masked else None
    target = torch.arange(B) % 2
    _, loss_ref, grads = oracle_full_train_step(sd, x, mask, target)
    m = DinoV2ClassifierSlice(1, 2, pretrained=False, precision="bf16").cuda()
    m.load_state_dict(sd)
    m.train()
    opt = m.configure_optimizers()[0]
    opt.zero_grad()
    batch = {"source": x, "target": target.cuda()}
    if mask is not None:
        batch["src_key_padding_mask"] = mask
    loss = m.training_step(batch, 0)
    assert abs(float(loss.detach()) - float(loss_ref)) <= 2e-2
    loss.backward()
    bad = {}
    for n, p in m.named_parameters():
        if n == "encoder.mask_token" or n.endswith("attn.qkv.bias") or n.endswith("in_proj_bias"):
            continue
        want, got = grads[n], p.grad.detach().cpu()
        if float(want.norm()) < 1e-12:
            assert float(got.norm()) < 1e-6, n
            continue
        cos, ratio = _cosine(got, want), float(got.norm()) / float(want.norm())
        if not (cos >= 0.99 and 0.95 <= ratio <= 1.05):
            bad[n] = (cos, ratio)
    assert not bad, bad
    opt.step()                           # a second step on the updated weights stays finite
    opt.zero_grad()
    assert torch.isfinite(m.training_step(batch, 1).detach())
