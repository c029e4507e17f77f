"""Generate the golden vectors under tests/golden/ by running the REAL reference model.

Run in the dev container only (needs /root/reference):   python tests/golden/make_golden.py
The reference is imported through oracle/ref_harness.py (stub modules for the packages that
are missing here).  Weights and inputs come from new-vit_b200/synth.py, which is deterministic,
so the GPU box can rebuild the identical state_dict / volumes and compare against these files.

Per case we store what the hot path produces:
  logits, logits_nosave (save_attn=False -> SDPA path), feat (without_linear), enc_cls,
  plane_cls = attention_maps[-1][:,:,0,:]   (captured before the getters mutate it in place),
  slice_cls = attention_maps_slice[-1][:,:,0,:],
  attn_maps = get_attention_maps(), slice_attn = get_slice_attention(),
  sal_sub   = run_pred()'s trilinear-upsampled map (scripts/main_predict.py:70-105,161-162),
              computed per volume (the script hard-codes batch 1) and subsampled [::7, ::7].
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_harness import build_reference_model  # noqa: E402
import new_vit_b200.synth as synth  # noqa: E402

CASES = {
    # name: (model_size, variant, B, D, H, W, masked, weight seed, volume seed)
    "s_init_b2": ("s", "init", 2, 32, 224, 224, False, 0, 0),
    "s_peaky_mask_b2": ("s", "peaky", 2, 32, 224, 224, True, 1, 1),
    "s_init_small": ("s", "init", 1, 5, 112, 112, False, 2, 2),
    "s_peaky_small_mask_b3": ("s", "peaky", 3, 7, 112, 112, True, 3, 3),
    # config 4 of BASELINE.json: ViT-B/14 encoder at 252x252 (256x256 is rejected by the reference, patch_embed.py:72-73)
    "b_peaky_252_mask_b2": ("b", "peaky", 2, 3, 252, 252, True, 4, 4),
    # ... and at config 4's full slice count: D = 64 => slice-transformer L = 65 with 12 heads x 64 (E = 768)
    "b_peaky_252_d64_mask_b1": ("b", "peaky", 1, 64, 252, 252, True, 13, 13),
    # hub-checkpoint layout: LayerScale gammas + "encoder.blocks.<i>" key names
    "s_hub_layerscale_b1": ("s", "peaky", 1, 4, 224, 224, False, 5, 5, {"hub_layout": True}),
    # ---- SURVEY 8f.2/8f.3 rows (flags: see run_case) ----
    # the hub "_reg" architecture: 4 register tokens, LayerScale, pos_embed built for 518 (37x37) resampled to 16x16
    "s_hub_reg518_b1": ("s", "peaky", 1, 3, 224, 224, False, 6, 6, {"hub_layout": True, "num_registers": 4, "pos_img": 518}),
    # local factory (pos_embed 16x16) on a non-square 126x168 input: bicubic resampling to 9x12
    "s_interp_126x168_b2": ("s", "peaky", 2, 3, 126, 168, True, 7, 7, {"pos_img": 224}),
    "s_bottleneck_posemb_b2": ("s", "peaky", 2, 5, 112, 112, True, 8, 8, {"use_bottleneck": True, "use_slice_pos_emb": True}),
    # rotary_positional_encoding='RoPE' on the slice tokens, with the bottleneck (the combination the reference's own
    # tests/models/test_dinoslice.py constructs) and a padding mask
    "s_rope_bottleneck_mask_b2": ("s", "peaky", 2, 9, 112, 112, True, 11, 11, {"use_bottleneck": True, "rope": True}),
    "s_rope_b2": ("s", "init", 2, 32, 112, 112, False, 12, 12, {"rope": True}),
    # rotary_positional_encoding='LiRE': runs in the reference for batch 1 and 32 slices only (it raises otherwise)
    "s_liere_mask_b1": ("s", "peaky", 1, 32, 112, 112, False, 14, 14, {"liere": True, "mask_tail": 5}),
    "s_fusion_linear_b2": ("s", "init", 2, 32, 56, 56, False, 9, 9, {"slice_fusion": "linear"}),
    "s_fusion_average_nolinear_b2": ("s", "init", 2, 6, 56, 56, False, 10, 10, {"slice_fusion": "average", "enable_linear": False}),
}


def run_case(name):
    size, variant, B, D, H, W, masked, wseed, vseed = CASES[name][:9]
    flags = CASES[name][9] if len(CASES[name]) > 9 else {}
    hub = bool(flags.get("hub_layout", False))
    nreg = int(flags.get("num_registers", 0))
    pos_img = int(flags.get("pos_img", H))
    fusion = flags.get("slice_fusion", "transformer")
    ctor = {k: flags[k] for k in ("use_bottleneck", "use_slice_pos_emb", "slice_fusion", "enable_linear") if k in flags}
    rope = bool(flags.get("rope", False))
    liere = bool(flags.get("liere", False))
    sd = synth.make_state_dict(size, out_ch=2, seed=wseed, variant=variant, img_size=pos_img, layerscale=hub,
                               chunked_names=not hub, num_registers=nreg, rope=rope, liere=liere, **ctor)
    ref_kw = dict(ctor, rotary_positional_encoding="RoPE") if rope else (dict(ctor, rotary_positional_encoding="LiRE") if liere else ctor)
    model = build_reference_model(sd, out_ch=2, model_size=size, hub_layout=hub, num_registers=nreg,
                                  pos_img_size=pos_img if pos_img != H or nreg else None, **ref_kw)
    x = synth.make_volume(B, D, H, W, seed=vseed)
    mask = synth.make_padding_mask(B, D, seed=vseed) if masked else None
    if flags.get("mask_tail"):   # batch 1: make_padding_mask never masks volume 0
        mask = torch.zeros(B, D, dtype=torch.bool)
        mask[:, D - flags["mask_tail"]:] = True
        masked = True
    out = {}
    with torch.no_grad():
        out["logits_nosave"] = model(x, src_key_padding_mask=mask, save_attn=False)
        out["feat"] = model(x, src_key_padding_mask=mask, save_attn=False, without_linear=True)
        xs = x.permute(0, 2, 1, 3, 4).reshape(B * D, H, W)[:, None].repeat(1, 3, 1, 1)
        out["enc_cls"] = model.encoder(xs)
        if pos_img != H or H != W:
            out["pos_embed"] = model.encoder.interpolate_pos_encoding(torch.zeros(1, 1 + (H // 14) * (W // 14), sd["encoder.pos_embed"].shape[-1]), H, W)
        if fusion == "transformer":
            out["logits"] = model(x, src_key_padding_mask=mask, save_attn=True)
            out["plane_cls"] = model.attention_maps[-1][:, :, 0, :].clone()
            out["slice_cls"] = model.attention_maps_slice[-1][:, :, 0, :].clone()
            if "rollout" in flags or nreg or name == "s_init_small":
                out["rollout_cls"] = model.get_attention_cls()[:, :, 0, :].clone()   # row 0 of the rollout product
            out["attn_maps"] = model.get_attention_maps().clone()
            out["slice_attn"] = model.get_slice_attention().clone()
            # saliency exactly as scripts/main_predict.py does it, one volume at a time
            subs = []
            for b in range(B):
                xb = x[b:b + 1]
                mb = None if mask is None else mask[b:b + 1]
                model(xb, src_key_padding_mask=mb, save_attn=True)
                w = model.get_attention_maps()
                w = w.mean(dim=1)
                if H == W:
                    g = int(w.shape[-1] ** 0.5)           # main_predict.py:93-94 assumes a square grid
                    w = w.view(1, 1, D, g, g)
                else:
                    w = w.view(1, 1, D, H // 14, W // 14)
                w = F.interpolate(w, size=xb.shape[2:], mode="trilinear")
                subs.append(w[0, 0, :, ::7, ::7].clone())
                if b == 0:
                    out["sal_sum_b0"] = w.double().sum().float().reshape(1)
                    out["sal_quantiles_b0"] = torch.from_numpy(np.quantile(w.numpy(), [0.5, 0.995, 0.999]))
            out["sal_sub"] = torch.stack(subs)
    meta = dict(size=size, variant=variant, B=B, D=D, H=H, W=W, masked=masked, wseed=wseed, vseed=vseed, hub_layout=hub,
                num_registers=nreg, pos_img=pos_img, mask_tail=int(flags.get("mask_tail", 0)),
                **(dict(ctor, rope=True) if rope else (dict(ctor, liere=True) if liere else ctor)))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"),
                        meta=np.array(repr(meta)), **{k: v.numpy() for k, v in out.items()})
    print(name, "logits", out["logits_nosave"].tolist()[:2])


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    for n in (sys.argv[1:] or CASES):
        run_case(n)
