"""CPU-side checks of the training surface: the oracle's slice transformer under torch autograd reproduces the gradients the REAL
reference produced (tests/golden/train_*.npz, made by tests/golden/make_train_golden.py), the Lightning-module methods of
base_model.py exist with the reference's semantics, and rotary_positional_encoding='LiRE' mirrors what the reference does with it
(it raises)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from new_vit_b200 import DinoV2ClassifierSlice, synth
from new_vit_b200.training import SLICE_PARAM_NAMES
from oracle import mst_oracle as O
from oracle import ref_harness


def oracle_train_step(sd, x, mask, target):
    """The reference's training step restated on the oracle: frozen encoder (no grad), slice transformer + head under autograd."""
    enc = O.forward(sd, x, mask)["enc_cls"].reshape(x.shape[0], x.shape[2], -1)
    params = {n: sd[n].clone().requires_grad_(True) for n in SLICE_PARAM_NAMES}
    full = dict(sd)
    full.update(params)
    B = x.shape[0]
    tok = torch.cat([params["cls_token"].repeat(B, 1, 1), enc], dim=1)
    kpm = None if mask is None else torch.cat([torch.zeros((B, 1), dtype=torch.bool), mask.bool()], dim=1)
    y, _ = O.slice_transformer(full, tok, kpm)
    logits = F.linear(y[:, 0], params["linear.weight"], params["linear.bias"])
    loss = F.cross_entropy(logits, target)
    loss.backward()
    return logits.detach(), loss.detach(), {n: p.grad for n, p in params.items()}


@pytest.mark.parametrize("name", ["train_s_frozen_b3", "train_s_frozen_d32_b2"])
def test_oracle_autograd_matches_reference_gradients(name):
    meta, g = load_golden(name)
    sd = synth.make_state_dict("s", 2, seed=meta["wseed"], variant="peaky", img_size=meta["H"])
    x = synth.make_volume(meta["B"], meta["D"], meta["H"], meta["W"], seed=meta["vseed"])
    mask = synth.make_padding_mask(meta["B"], meta["D"], seed=meta["vseed"]) if meta["masked"] else None
    assert sorted(meta["trainable"]) == sorted(SLICE_PARAM_NAMES)    # what freeze=True leaves trainable (dino.py:69-71)
    logits, loss, grads = oracle_train_step(sd, x, mask, g["target"])
    torch.testing.assert_close(logits, g["logits"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(loss.reshape(1), g["loss"], rtol=1e-5, atol=1e-6)
    sub = meta["sub"]
    for n in SLICE_PARAM_NAMES:
        want = g["grad." + n]
        got = grads[n][::sub] if grads[n].dim() == 2 and grads[n].shape[0] >= 64 else grads[n]
        torch.testing.assert_close(got, want, rtol=1e-3, atol=1e-7, msg=lambda m, n=n: f"{n}: {m}")


def test_lightning_surface_mirrors_base_model():
    m = DinoV2ClassifierSlice(in_ch=1, out_ch=2, pretrained=False, freeze=True)
    for name in ("training_step", "validation_step", "test_step", "_step", "_epoch_end", "compute_loss", "configure_optimizers",
                 "on_train_epoch_end", "on_validation_epoch_end", "on_test_epoch_end", "log"):
        assert callable(getattr(m, name)), name
    assert m.optimizer is torch.optim.AdamW and m.optimizer_kwargs == {'lr': 1e-6, 'weight_decay': 1e-2}      # dino.py:41
    assert isinstance(m.loss_func, torch.nn.CrossEntropyLoss)                                                  # base_model.py:124,139
    assert all(not p.requires_grad for p in m.encoder.parameters())                                             # dino.py:69-71
    trainable = [n for n, p in m.named_parameters() if p.requires_grad]
    assert sorted(trainable) == sorted(SLICE_PARAM_NAMES)
    opt = m.configure_optimizers()                                     # on the CPU: the reference's own optimizer class
    assert isinstance(opt, list) and isinstance(opt[0], torch.optim.AdamW)
    pred, target = torch.tensor([[2.0, 0.0], [0.0, 1.0]]), torch.tensor([0, 0])
    assert torch.equal(m.compute_loss(pred, target), F.cross_entropy(pred, target))
    m.acc["val_"].update(pred, target)
    m.auc_roc["val_"].update(pred, torch.tensor([0, 1]))
    m.batch_size = 2
    m._epoch_end("val")
    assert float(m.logged["val/ACC"]) == 0.5 and float(m.logged["val/AUC_ROC"]) == 1.0


def test_unfrozen_training_forward_says_what_is_missing():
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the GPU tests")
    m = DinoV2ClassifierSlice(in_ch=1, out_ch=2, pretrained=False).train()
    with pytest.raises(Exception):      # no CUDA device here: MSTError; on a GPU: NotImplementedError (encoder backward not built)
        m(torch.zeros(1, 1, 2, 28, 28))


def test_liere_mirrors_the_reference():
    """rotary_positional_encoding='LiRE' (transformer_blocks.py:352-358): same extra state_dict tensors; the reference evaluates it
    for batch 1 and 32 slices only and raises RuntimeError otherwise (rotary_embedding_torch.py:350 hard-codes 33 tokens; the
    permuted result cannot be viewed as [B*heads, L, hd] for B > 1, transformer_blocks.py:263) -- same errors, same messages here.
    (What it computes for batch 1 is pinned by the golden s_liere_mask_b1: oracle on the CPU, CUDA kernel in the GPU tests.)"""
    ours = DinoV2ClassifierSlice(in_ch=1, out_ch=2, pretrained=False, rotary_positional_encoding='LiRE')
    keys = [k for k in ours.state_dict() if "rotary" in k]
    assert keys == [f"slice_fusion.layers.0.self_attn.rotary_positional_encoding.vars.{i}" for i in range(2)]
    assert all(tuple(ours.state_dict()[k].shape) == (120, 33, 1) for k in keys) and len(ours.state_dict()) == 170
    errs = {}
    for B, D in ((2, 32), (1, 5)):
        with pytest.raises(RuntimeError) as e:
            ours(torch.zeros(B, 1, D, 56, 56))
        errs[(B, D)] = str(e.value)
    assert "view size is not compatible" in errs[(2, 32)] and "shape '[1, 33, 12, 32]' is invalid for input of size 2304" in errs[(1, 5)]
    if ref_harness.reference_available():
        Ref = ref_harness.load_reference_class()
        ref = Ref(in_ch=1, out_ch=2, pretrained=False, rotary_positional_encoding='LiRE').eval()
        assert list(ref.state_dict().keys()) == list(ours.state_dict().keys())
        ref.load_state_dict(ours.state_dict())
        for B, D in ((2, 32), (1, 5)):
            with pytest.raises(RuntimeError) as e:
                ref(torch.zeros(B, 1, D, 56, 56))
            assert str(e.value) == errs[(B, D)]
        with torch.no_grad():
            assert ref(torch.zeros(1, 1, 32, 56, 56)).shape == (1, 2)     # the one shape it evaluates


def oracle_full_train_step(sd, x, mask, target):
    """Every parameter trainable (main_train.py:110-126 on the default construction): the oracle's functions under torch autograd."""
    params = {k: v.clone().float().requires_grad_(True) for k, v in sd.items() if k != "encoder.mask_token"}
    full = dict(sd)
    full.update(params)
    B, C, D, H, W = x.shape
    img = x.float().permute(0, 2, 1, 3, 4).reshape(B * D, H, W)[:, None].repeat(1, 3, 1, 1)
    enc, _ = O.encoder_forward(full, img, params["encoder.pos_embed"].shape[-1] // 64)
    tok = torch.cat([params["cls_token"].repeat(B, 1, 1), enc.reshape(B, D, -1)], dim=1)
    kpm = None if mask is None else torch.cat([torch.zeros((B, 1), dtype=torch.bool), mask.bool()], dim=1)
    y, _ = O.slice_transformer(full, tok, kpm)
    logits = F.linear(y[:, 0], params["linear.weight"], params["linear.bias"])
    loss = F.cross_entropy(logits, target)
    loss.backward()
    return logits.detach(), loss.detach(), {n: p.grad for n, p in params.items()}


def test_oracle_autograd_matches_reference_gradients_full_model():
    meta, g = load_golden("train_s_full_b2")
    sd = synth.make_state_dict("s", 2, seed=meta["wseed"], variant="peaky", img_size=meta["H"])
    x = synth.make_volume(meta["B"], meta["D"], meta["H"], meta["W"], seed=meta["vseed"])
    mask = synth.make_padding_mask(meta["B"], meta["D"], seed=meta["vseed"])
    logits, loss, grads = oracle_full_train_step(sd, x, mask, g["target"])
    torch.testing.assert_close(loss.reshape(1), g["loss"], rtol=1e-5, atol=1e-6)
    assert len(meta["trainable"]) == 167           # all 168 tensors but encoder.mask_token (never read)
    for n in meta["trainable"]:
        got = grads[n].reshape(-1)
        torch.testing.assert_close(got[::meta["stride"]], g["grad." + n], rtol=2e-3, atol=2e-5 * float(g["norm." + n]) + 1e-10,
                                   msg=lambda m, n=n: f"{n}: {m}")
