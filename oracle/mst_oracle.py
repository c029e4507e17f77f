"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (PyTorch fp32) restatement of the reference's MST-DINOv2 hot path, written as plain
functions over a state_dict so that it can run on the GPU box where `/root/reference` does not
exist.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this module; the product package (`new-vit_b200/`) never does.

Pinning: the reference holds no golden vectors or asserting tests for this path (SURVEY.md
section 4 -- "parity unpinned" by the reference's own tests).  The oracle is therefore pinned
against outputs of the reference itself, run in the dev container through
`oracle/ref_harness.py`: `tests/golden/make_golden.py` committed `tests/golden/*.npz`, and
`tests/test_oracle.py` checks this file against them (and, when `/root/reference` is present,
against the live reference model).

Every function cites the reference file:line it restates (paths relative to /root/reference).
"""
import math

import torch
import torch.nn.functional as F

PATCH = 14


def _blk_prefix(sd, i):
    # local model: encoder.blocks.0.<i>. (BlockChunk wrapper, vision_transformer.py:153-160);
    # hub model: encoder.blocks.<i>.
    p = f"encoder.blocks.0.{i}."
    if p + "norm1.weight" in sd:
        return p
    return f"encoder.blocks.{i}."


def encoder_depth(sd):
    d = 0
    while _blk_prefix(sd, d) + "norm1.weight" in sd:
        d += 1
    return d


def interpolate_pos_encoding(pos_embed, npatch, w, h, interpolate_offset=0.1, interpolate_antialias=False):
    """vision_transformer.py:179-211.  (interpolate_offset, interpolate_antialias) = (0.1, False) for the vendored factory
    and the plain hub checkpoints, (0.0, True) for the hub "_reg" checkpoints (dino.py:60-61)."""
    N = pos_embed.shape[1] - 1
    if npatch == N and w == h:
        return pos_embed
    pos_embed = pos_embed.float()
    class_pos = pos_embed[:, 0]
    patch_pos = pos_embed[:, 1:]
    dim = pos_embed.shape[-1]
    w0, h0 = w // PATCH, h // PATCH
    M = int(math.sqrt(N))
    assert N == M * M
    kwargs = {}
    if interpolate_offset:
        kwargs["scale_factor"] = (float(w0 + interpolate_offset) / M, float(h0 + interpolate_offset) / M)
    else:
        kwargs["size"] = (w0, h0)
    patch_pos = F.interpolate(patch_pos.reshape(1, M, M, dim).permute(0, 3, 1, 2), mode="bicubic",
                              antialias=interpolate_antialias, **kwargs)
    assert (w0, h0) == patch_pos.shape[-2:]
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((class_pos.unsqueeze(0), patch_pos), dim=1)


def encoder_attention(sd, p, x, heads):
    """layers/attention.py:56-69 (the explicit-softmax path; xFormers is not installed)."""
    B, N, C = x.shape
    qkv = F.linear(x, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    scale = (C // heads) ** -0.5
    q, k, v = qkv[0] * scale, qkv[1], qkv[2]          # scale into q BEFORE q@k^T (attention.py:60)
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)  # [B, heads, N, N]
    o = (attn @ v).transpose(1, 2).reshape(B, N, C)
    o = F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    return o, attn


def encoder_forward(sd, img, heads, keep_all_maps=False):
    """DinoVisionTransformer.forward (vision_transformer.py:324-329) on `[BD,3,H,W]`.

    Returns normed CLS `[BD,E]` and the last block's attention `[BD,heads,N,N]` (or all 12).
    """
    BD, _, H, W = img.shape
    assert H % PATCH == 0 and W % PATCH == 0, "patch_embed.py:72-73"
    # patch_embed.py:75-77: Conv2d(k=s=14) -> flatten(2).transpose(1,2)
    x = F.conv2d(img, sd["encoder.patch_embed.proj.weight"], sd["encoder.patch_embed.proj.bias"], stride=PATCH)
    x = x.flatten(2).transpose(1, 2)
    # vision_transformer.py:219-220
    x = torch.cat((sd["encoder.cls_token"].expand(BD, -1, -1), x), dim=1)
    # (the reference names the image dims "w, h = x.shape[2:]": its w is the height)
    reg = "encoder.register_tokens" in sd   # the "_reg" hub architecture resamples anti-aliased, without the 0.1 offset
    x = x + interpolate_pos_encoding(sd["encoder.pos_embed"], x.shape[1] - 1, H, W, 0.0 if reg else 0.1, reg)
    if "encoder.register_tokens" in sd:  # vision_transformer.py:222-230: registers go in AFTER the position add
        x = torch.cat((x[:, :1], sd["encoder.register_tokens"].expand(BD, -1, -1), x[:, 1:]), dim=1)
    maps = []
    depth = encoder_depth(sd)
    for i in range(depth):
        p = _blk_prefix(sd, i)
        # block.py:112-113, LayerNorm eps 1e-6 (vision_transformer.py:95), LayerScale (layer_scale.py:26-27)
        a, attn = encoder_attention(sd, p, F.layer_norm(x, x.shape[-1:], sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6), heads)
        if p + "ls1.gamma" in sd:
            a = a * sd[p + "ls1.gamma"]
        x = x + a
        h = F.layer_norm(x, x.shape[-1:], sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
        h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
        h = F.gelu(h)  # exact erf GELU (mlp.py:30)
        h = F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        if p + "ls2.gamma" in sd:
            h = h * sd[p + "ls2.gamma"]
        x = x + h
        if keep_all_maps or i == depth - 1:
            maps.append(attn)
    x = F.layer_norm(x, x.shape[-1:], sd["encoder.norm.weight"], sd["encoder.norm.bias"], 1e-6)
    return x[:, 0], maps


def rope_rotate(t, freqs):
    """RotaryEmbedding.rotate_queries_or_keys (utils/rotary_embedding_torch.py:159-173,45-62,38-42,272-300) for the
    configuration the slice transformer builds (transformer_blocks.py:335-351: dim = head_dim, theta = 256, 'lang' frequencies
    1/theta^(2i/dim), no xpos, interpolate_factor 1): token at sequence position j (0 = the slice-CLS token), feature pair
    (2i, 2i+1) rotated by the angle j*freqs[i].  t: [B, heads, L, hd]."""
    L = t.shape[-2]
    ang = torch.arange(L, dtype=t.dtype)[:, None] * freqs[None, :]          # einsum('..., f -> ... f', seq, freqs)
    ang = ang.repeat_interleave(2, dim=-1)                                   # repeat '... n -> ... (n r)', r = 2
    t1, t2 = t[..., 0::2], t[..., 1::2]
    rot = torch.stack((-t2, t1), dim=-1).reshape(t.shape)                    # rotate_half
    return t * ang.cos() + rot * ang.sin()


def liere_rotate(t, liere_vars):
    """rotary_positional_encoding='LiRE' as the reference applies it to q and k (transformer_blocks.py:263-264 ->
    AttentionLiereRotator.rotate_queries_or_keys / forward, rotary_embedding_torch.py:346-396), operation by operation -- including
    the two places where it can only raise: the hard-coded 33 tokens (:350) and the caller's .view() of the permuted result
    (transformer_blocks.py:263), which succeeds for batch 1 only.  t: [B, heads, L, hd] -> [B*heads, L, hd].

    What survives for B = 1, L = 33: ONE position-independent orthogonal matrix R (matrix_exp of a skew matrix built from all 33
    generator slices) multiplies every vector, and the final view re-reads the [hd, L, heads] result as [heads, L, hd], so slot
    (head i, position j) receives the vector of position (33 i + j) // heads, head (33 i + j) % heads.  R cancels in q . k."""
    B, H, L, hd = t.shape
    blk = hd // len(liere_vars)
    t = t.permute(0, 2, 1, 3).contiguous()                                    # :392
    x = t.view(B, 33 * 1, H, hd)                                              # :350 (axes_length = 33, spacial_dims = 1)
    mats = []
    for v in liere_vars:                                                      # :359-362, flat_to_skew :320-327
        A = torch.zeros(blk, blk, 33, 1)
        i, j = torch.tril_indices(blk, blk, offset=-1)
        A[i, j, :, 0] = v[:, :, 0]
        A[j, i, :, 0] = -v[:, :, 0]
        mats.append(A.view(blk, blk, 33) @ torch.arange(0, 33).to(t.dtype))
    R = torch.block_diag(*[torch.linalg.matrix_exp(A.float()) for A in mats])  # :365,371
    x = x.permute(0, 3, 1, 2)                                                 # :381
    y = torch.bmm(R.unsqueeze(0).repeat(B, 1, 1), x.reshape(B, hd, 33 * H).float())   # :386 (the reference's sparse bmm)
    y = y.view(B, hd, 33, H).permute(0, 2, 3, 1)                              # :387
    return y.view(B * H, L, hd)                                               # transformer_blocks.py:263


def slice_transformer(sd, x, key_padding_mask, heads=12):  # noqa: C901
    """nn.TransformerEncoder(num_layers=1, norm) over the custom pre-LN layer
    (utils/transformer_blocks.py:524-573, 29-318; dino.py:84-96).  Explicit bmm/softmax path
    (`:266-295`), which the reference takes when save_attn=True; the SDPA path differs by ~2e-7.
    x: [B,L,E]; key_padding_mask: bool [B,L] (True = ignore) or None.
    Returns [B,L,E] and per-head weights [B,heads,L,L]."""
    q_ = "slice_fusion.layers.0."
    B, L, E = x.shape
    hd = E // heads
    h = F.layer_norm(x, (E,), sd[q_ + "norm1.weight"], sd[q_ + "norm1.bias"], 1e-5)
    qkv = F.linear(h, sd[q_ + "self_attn.in_proj_weight"], sd[q_ + "self_attn.in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)
    q = q.reshape(B, L, heads, hd).transpose(1, 2) * math.sqrt(1.0 / hd)  # transformer_blocks.py:268
    k = k.reshape(B, L, heads, hd).transpose(1, 2)
    v = v.reshape(B, L, heads, hd).transpose(1, 2)
    rope = q_ + "self_attn.rotary_positional_encoding.freqs"
    if rope in sd:  # rotary_positional_encoding='RoPE' (transformer_blocks.py:262-264,335-351): q and k rotated by position
        q, k = rope_rotate(q, sd[rope]), rope_rotate(k, sd[rope])
    lv = q_ + "self_attn.rotary_positional_encoding.vars."
    if lv + "0" in sd:   # rotary_positional_encoding='LiRE'
        liere_vars = [sd[lv + str(i)] for i in range(2)]
        q = liere_rotate(q, liere_vars).view(B, heads, L, hd)
        k = liere_rotate(k, liere_vars).view(B, heads, L, hd)
    s = q @ k.transpose(-2, -1)
    if key_padding_mask is not None:  # additive -inf on key columns (transformer_blocks.py:244-252)
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    w = s.softmax(dim=-1)
    o = (w @ v).transpose(1, 2).reshape(B, L, E)
    o = F.linear(o, sd[q_ + "self_attn.out_proj.weight"], sd[q_ + "self_attn.out_proj.bias"])
    x = x + o
    h = F.layer_norm(x, (E,), sd[q_ + "norm2.weight"], sd[q_ + "norm2.bias"], 1e-5)
    h = F.linear(F.relu(F.linear(h, sd[q_ + "linear1.weight"], sd[q_ + "linear1.bias"])),
                 sd[q_ + "linear2.weight"], sd[q_ + "linear2.bias"])  # ReLU FFN (transformer_blocks.py:484,585)
    x = x + h
    x = F.layer_norm(x, (E,), sd["slice_fusion.norm.weight"], sd["slice_fusion.norm.bias"], 1e-5)  # dino.py:95
    return x, w


@torch.no_grad()
def forward(sd, source, src_key_padding_mask=None, enc_heads=None, keep_all_maps=False, slice_fusion="transformer"):
    """DinoV2ClassifierSlice.forward (dino.py:110-167).  The constructor variants are read off the state_dict
    (bottleneck.*, slice_pos_emb.weight, encoder.register_tokens, linear.* present or not); `slice_fusion`
    selects dino.py:144-157; RoPE / LiRE on the slice tokens are applied when their `freqs` / `vars.*` tensors are in the
    state_dict.

    Returns a dict: logits [B,out] (None without the linear head), feat, enc_cls [BD,E], plane_cls [BD,heads,N]
    (row 0 of the last encoder block's attention), slice_cls [B,12,L] (row 0 of slice attention), and the
    full maps under 'maps' / 'maps_slice' as the reference stores them (dino.py:241,249)."""
    sd = {k: v.float() for k, v in sd.items()}
    B, C, D, H, W = source.shape
    assert C == 1
    E = sd["encoder.pos_embed"].shape[-1]
    if enc_heads is None:
        enc_heads = E // 64
    x = source.float().permute(0, 2, 1, 3, 4).reshape(B * D * C, H, W)  # dino.py:125
    x = x[:, None].repeat(1, 3, 1, 1)                                    # dino.py:126-127
    enc_cls, maps = encoder_forward(sd, x, enc_heads, keep_all_maps)     # dino.py:131
    x = enc_cls
    if "bottleneck.weight" in sd:                                        # dino.py:134-135
        x = F.linear(x, sd["bottleneck.weight"], sd["bottleneck.bias"])
    x = x.reshape(B, D, -1)                                              # dino.py:138
    if "slice_pos_emb.weight" in sd:                                     # dino.py:140-142
        x = x + sd["slice_pos_emb.weight"][:D]
    if slice_fusion != "transformer":
        feat = x.reshape(B, -1) if slice_fusion == "linear" else x.mean(dim=1)   # dino.py:154-157
        logits = F.linear(feat, sd["linear.weight"], sd["linear.bias"]) if "linear.weight" in sd else None
        return {"logits": logits, "feat": feat, "enc_cls": enc_cls, "plane_cls": maps[-1][:, :, 0, :].clone(),
                "slice_cls": None, "maps": maps, "maps_slice": []}
    x = torch.cat([sd["cls_token"].repeat(B, 1, 1), x], dim=1)           # dino.py:145
    kpm = None
    if src_key_padding_mask is not None:                                 # dino.py:147-150
        kpm = torch.cat([torch.zeros((B, 1), dtype=torch.bool), src_key_padding_mask.bool()], dim=1)
    x, w = slice_transformer(sd, x, kpm)                                 # dino.py:152
    feat = x[:, 0]                                                       # dino.py:153
    logits = F.linear(feat, sd["linear.weight"], sd["linear.bias"]) if "linear.weight" in sd else None  # dino.py:166,103
    return {
        "logits": logits, "feat": feat, "enc_cls": enc_cls,
        "plane_cls": maps[-1][:, :, 0, :].clone(), "slice_cls": w[:, :, 0, :].clone(),
        "maps": maps, "maps_slice": [w],
    }


def get_plane_attention(plane_cls, num_registers=0):
    """dino.py:189-195 on the CLS row [BD,heads,N]: drop CLS (+ register) columns, zero patch 0, renormalise."""
    a = plane_cls[:, :, 1 + num_registers:].clone()
    a[:, :, 0] = 0
    a = a / a.sum(dim=-1, keepdim=True)
    return a


def get_slice_attention(slice_cls):
    """dino.py:173-187 on the CLS row [B,heads,L] -> [B*D,1,1]."""
    s = slice_cls[:, :, 1:].clone()
    s = s / s.sum(dim=-1, keepdim=True)
    s = s.mean(dim=1)
    return s.reshape(-1)[:, None, None]


def get_attention_maps(plane_cls, slice_cls, num_registers=0):
    """dino.py:197-202 -> [BD,heads,P]."""
    return get_slice_attention(slice_cls) * get_plane_attention(plane_cls, num_registers)


def quantile(x, q):
    """np.quantile as scripts/main_predict.py:243-245,296 calls it (default 'linear' method), per item of [items, ...]."""
    import numpy as np
    a = x.detach().cpu().numpy().reshape(x.shape[0], -1)
    return torch.from_numpy(np.stack([np.quantile(r, list(q)) for r in a]))


def get_attention_cls(maps):
    """dino.py:204-212 attention rollout over all encoder blocks."""
    a = maps[-1]
    for m in reversed(maps[:-1]):
        a = torch.matmul(m, a)
    return a


def saliency(plane_cls, slice_cls, B, D, H, W, num_registers=0):
    """scripts/main_predict.py:70-105,161-162 generalised from B=1 to a batch:
    head-mean of get_attention_maps -> [B,1,D,g,g] -> trilinear upsample to [B,1,D,H,W];
    slice weights broadcast to the source shape.  Returns (coarse, full, weight_slice)."""
    w = get_attention_maps(plane_cls, slice_cls, num_registers).mean(dim=1)  # main_predict.py:73-74
    if H == W:
        g = int(w.shape[-1] ** 0.5)                            # main_predict.py:93-94 (assumes a square patch grid)
        coarse = w.reshape(B, 1, D, g, g)                      # main_predict.py:100 (B=1 there)
    else:
        coarse = w.reshape(B, 1, D, H // PATCH, W // PATCH)
    full = F.interpolate(coarse, size=(D, H, W), mode="trilinear")  # main_predict.py:161-162
    ws = get_slice_attention(slice_cls).mean(dim=1)            # main_predict.py:103-104
    ws = ws.reshape(B, 1, D, 1, 1).expand(B, 1, D, H, W)
    return coarse, full, ws


def bilinear14_reference(coarse, H, W):
    """with depth scale 1 the trilinear upsample equals per-slice bilinear with
    align_corners=False; plain-loop restatement used to pin the CUDA kernel's index math."""
    B, _, D, gh, gw = coarse.shape
    sy_scale, sx_scale = gh / H, gw / W
    ys = ((torch.arange(H, dtype=torch.float32) + 0.5) * sy_scale - 0.5).clamp(min=0)
    xs = ((torch.arange(W, dtype=torch.float32) + 0.5) * sx_scale - 0.5).clamp(min=0)
    y0 = ys.floor().long(); x0 = xs.floor().long()
    y1 = (y0 + 1).clamp(max=gh - 1); x1 = (x0 + 1).clamp(max=gw - 1)
    ly = (ys - y0).view(1, 1, 1, H, 1); lx = (xs - x0).view(1, 1, 1, 1, W)
    v00 = coarse[..., y0, :][..., x0]; v01 = coarse[..., y0, :][..., x1]
    v10 = coarse[..., y1, :][..., x0]; v11 = coarse[..., y1, :][..., x1]
    out = (1 - ly) * ((1 - lx) * v00 + lx * v01) + ly * ((1 - lx) * v10 + lx * v11)
    return out


def run_pred(sd, source, src_key_padding_mask=None, use_softmax=True, use_tta=False):
    """scripts/main_predict.py:133-164 (`run_pred` with `_pred_trans`, save_attn=True), generalised to a batch:
    softmax of the logits, head-mean saliency map reshaped to [B,1,D,g,g], slice weights broadcast to the source
    shape, optional 8-flip TTA (the un-flipped padding mask is passed to every flip, as the script does), and the
    trilinear upsample AFTER the TTA average (:161-162).  Returns (pred, weight [B,1,D,H,W], weight_slice)."""
    B, _, D, H, W = source.shape

    def pred_trans(x):
        r = forward(sd, x, src_key_padding_mask)
        pred = torch.softmax(r["logits"], dim=-1) if use_softmax else r["logits"]
        w = get_attention_maps(r["plane_cls"], r["slice_cls"]).mean(dim=1)
        g = int(w.shape[-1] ** 0.5)
        w = w.reshape(B, 1, D, g, g)
        ws = get_slice_attention(r["slice_cls"]).mean(dim=1).reshape(B, 1, D, 1, 1) * torch.ones_like(x)
        return pred, w, ws

    pred, weight, weight_slice = pred_trans(source)
    if use_tta:
        for flip_dim in [(2,), (3,), (4,), (2, 3), (2, 4), (3, 4), (2, 3, 4)]:
            p_i, w_i, ws_i = pred_trans(torch.flip(source, flip_dim))
            pred = pred + p_i
            weight = weight + torch.flip(w_i, flip_dim)
            weight_slice = weight_slice + torch.flip(ws_i, flip_dim)
        pred, weight, weight_slice = pred / 8, weight / 8, weight_slice / 8
    weight = F.interpolate(weight, size=(D, H, W), mode="trilinear")
    return pred, weight, weight_slice
