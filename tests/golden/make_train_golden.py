"""Golden vectors for the training step (BASELINE.json config 5, frozen-encoder construction) from the REAL reference model:

    python tests/golden/make_train_golden.py          (dev container only: needs /root/reference)

`DinoV2ClassifierSlice(freeze=True)` in train() mode (dino.py:69-71; dropouts are 0, so train == eval arithmetic), one
`_step`-style pass: pred = model(source, src_key_padding_mask) -> CrossEntropyLoss(pred, target) (base_model.py:155-159,180-181)
-> loss.backward().  Stored: logits, loss, the gradient of every trainable tensor (the 17 slice-transformer / head tensors), and
the parameters after ONE torch.optim.AdamW step (base_model.py:103-110).  To keep the fixtures small the [E, E] / [3E, E]
matrices are stored as every 8th row (SUB); vectors are stored whole."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_harness import build_reference_model  # noqa: E402
import new_vit_b200.synth as synth  # noqa: E402

SUB = 8


def _sub(t):
    return t[::SUB] if t.dim() == 2 and t.shape[0] >= 64 else t


FULL_STRIDE = 61   # un-frozen case: every tensor flattened and stored as every 61st element (22.5 M gradients otherwise)

CASES = {
    # name: (B, D, H, W, masked, weight seed, volume seed, lr)
    "train_s_frozen_b3": (3, 8, 56, 56, True, 21, 21, 1e-3),
    "train_s_frozen_d32_b2": (2, 32, 28, 28, False, 22, 22, 1e-6),
    # the default construction of main_train.py: every parameter trains (224 x 224: the position table's own grid)
    "train_s_full_b2": (2, 4, 224, 224, True, 23, 23, 1e-6, "full"),
}


def run_case(name):
    B, D, H, W, masked, wseed, vseed, lr = CASES[name][:8]
    full = len(CASES[name]) > 8
    sd = synth.make_state_dict("s", out_ch=2, seed=wseed, variant="peaky", img_size=H)
    model = build_reference_model(sd, out_ch=2, model_size="s", pos_img_size=None if full else H, freeze=not full).train()
    for p in model.encoder.parameters():     # build_reference_model swaps the encoder in after __init__ froze the first one
        p.requires_grad = full
    x = synth.make_volume(B, D, H, W, seed=vseed)
    mask = synth.make_padding_mask(B, D, seed=vseed) if masked else None
    target = torch.tensor([(i * 7 + vseed) % 2 for i in range(B)])
    pred = model(x, src_key_padding_mask=mask)
    loss = torch.nn.CrossEntropyLoss()(pred, target)
    loss.backward()
    out = {"logits": pred.detach(), "loss": loss.detach().reshape(1), "target": target}
    names = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is not None]
    if full:   # strided subsample + norm of every gradient
        for n in names:
            gflat = dict(model.named_parameters())[n].grad.detach().reshape(-1)
            out["grad." + n] = gflat[::FULL_STRIDE].clone()
            out["norm." + n] = gflat.norm().reshape(1)
        meta = dict(B=B, D=D, H=H, W=W, masked=masked, wseed=wseed, vseed=vseed, lr=lr, trainable=names, stride=FULL_STRIDE, full=True)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), meta=np.array(repr(meta)),
                            **{k: v.numpy() for k, v in out.items()})
        print(name, "loss", float(loss), "trainable tensors with a gradient", len(names))
        return
    for n in names:
        out["grad." + n] = _sub(dict(model.named_parameters())[n].grad.detach().clone())
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=lr, weight_decay=1e-2)
    opt.step()
    for n in names:
        out["after." + n] = _sub(dict(model.named_parameters())[n].detach().clone())
    meta = dict(B=B, D=D, H=H, W=W, masked=masked, wseed=wseed, vseed=vseed, lr=lr, trainable=names, sub=SUB)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), meta=np.array(repr(meta)),
                        **{k: v.numpy() for k, v in out.items()})
    print(name, "loss", float(loss), "trainable tensors", len(names), "logits", pred.detach().tolist())


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    for n in (sys.argv[1:] or CASES):
        run_case(n)
