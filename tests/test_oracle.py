"""Pin the oracle (oracle/mst_oracle.py) against golden vectors produced by the real reference
(tests/golden/make_golden.py) and, when /root/reference is present, against the live model."""
import pytest
import torch

from conftest import case_inputs, load_golden
from oracle import mst_oracle as O
from oracle import ref_harness

SMALL = ["s_init_small", "s_peaky_small_mask_b3", "s_hub_layerscale_b1", "s_hub_reg518_b1", "s_interp_126x168_b2",
         "s_bottleneck_posemb_b2", "s_rope_bottleneck_mask_b2", "s_rope_b2", "s_liere_mask_b1"]
NO_TRANSFORMER = ["s_fusion_linear_b2", "s_fusion_average_nolinear_b2"]
FULL = ["s_init_b2", "s_peaky_mask_b2", "b_peaky_252_mask_b2", "b_peaky_252_d64_mask_b1"]


def _check(name):
    meta, g = load_golden(name)
    sd, x, mask = case_inputs(meta)
    nreg = meta.get("num_registers", 0)
    r = O.forward(sd, x, mask, keep_all_maps="rollout_cls" in g)
    B, D, H, W = meta["B"], meta["D"], meta["H"], meta["W"]
    if "pos_embed" in g:   # interpolate_pos_encoding (vision_transformer.py:179-211), bicubic through ATen
        torch.testing.assert_close(O.interpolate_pos_encoding(sd["encoder.pos_embed"], (H // 14) * (W // 14), H, W,
                                                              0.0 if nreg else 0.1, nreg > 0),
                                   g["pos_embed"], rtol=1e-5, atol=1e-6)
    if "rollout_cls" in g:
        torch.testing.assert_close(O.get_attention_cls(r["maps"])[:, :, 0, :], g["rollout_cls"], rtol=1e-4, atol=1e-8)
    # fp32 restatement of the same ATen ops: expect ~1e-6; tolerance 1e-5 abs / 1e-4 rel
    torch.testing.assert_close(r["logits"], g["logits"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(r["logits"], g["logits_nosave"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(r["feat"], g["feat"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(r["enc_cls"], g["enc_cls"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(r["plane_cls"], g["plane_cls"], rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(r["slice_cls"], g["slice_cls"], rtol=1e-4, atol=1e-7)
    maps = O.get_attention_maps(r["plane_cls"], r["slice_cls"], nreg)
    assert maps.shape == g["attn_maps"].shape
    torch.testing.assert_close(maps, g["attn_maps"], rtol=1e-4, atol=1e-9)
    torch.testing.assert_close(O.get_slice_attention(r["slice_cls"]), g["slice_attn"], rtol=1e-4, atol=1e-8)
    # argmax of the head-mean map: bit-exact indexing
    assert torch.equal(maps.mean(1).reshape(B, -1).argmax(-1), g["attn_maps"].mean(1).reshape(B, -1).argmax(-1))
    coarse, full, ws = O.saliency(r["plane_cls"], r["slice_cls"], B, D, H, W, nreg)
    if "sal_quantiles_b0" in g:   # np.quantile as main_predict.py:243-245,296 takes it
        torch.testing.assert_close(O.quantile(full[:1], [0.5, 0.995, 0.999])[0], g["sal_quantiles_b0"], rtol=1e-4, atol=0)
    torch.testing.assert_close(full[:, 0, :, ::7, ::7], g["sal_sub"], rtol=1e-4, atol=1e-10)
    torch.testing.assert_close(full[0].double().sum().float().reshape(1), g["sal_sum_b0"], rtol=1e-4, atol=0)
    # depth scale 1 => trilinear == per-slice bilinear: bit-identical between the two
    # ATen kernels; the plain-formula restatement differs only by FMA contraction (<= 2 ulp)
    bil = torch.nn.functional.interpolate(coarse[:, 0], size=(H, W), mode="bilinear", align_corners=False)
    assert torch.equal(bil[:, None], full)
    torch.testing.assert_close(O.bilinear14_reference(coarse, H, W), full, rtol=1e-5, atol=float(full.max()) * 1e-6)
    if meta["masked"]:
        m = case_inputs(meta)[2]
        assert (O.get_slice_attention(r["slice_cls"]).reshape(B, D)[m] == 0).all()


@pytest.mark.parametrize("name", SMALL)
def test_oracle_matches_reference_golden_small(name):
    _check(name)


@pytest.mark.parametrize("name", FULL)
def test_oracle_matches_reference_golden_full(name):
    _check(name)


@pytest.mark.parametrize("name", NO_TRANSFORMER)
def test_oracle_matches_reference_golden_other_fusions(name):
    """slice_fusion='linear' / 'average' (dino.py:154-157), with and without the linear head (dino.py:103)."""
    meta, g = load_golden(name)
    sd, x, mask = case_inputs(meta)
    r = O.forward(sd, x, mask, slice_fusion=meta["slice_fusion"])
    torch.testing.assert_close(r["feat"], g["feat"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(r["enc_cls"], g["enc_cls"], rtol=1e-4, atol=1e-5)
    if meta.get("enable_linear", True):
        torch.testing.assert_close(r["logits"], g["logits_nosave"], rtol=1e-4, atol=1e-5)
    else:
        assert r["logits"] is None
        torch.testing.assert_close(r["feat"], g["logits_nosave"], rtol=1e-4, atol=1e-5)   # Identity head returns the feature


@pytest.mark.skipif(not ref_harness.reference_available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference():
    meta, _ = load_golden("s_init_small")
    sd, x, mask = case_inputs(meta)
    model = ref_harness.build_reference_model(sd)
    with torch.no_grad():
        y = model(x, save_attn=True)
        # rollout first: the getters below mutate attention_maps[-1] in place
        ref_rollout = model.get_attention_cls().clone()
        ref_maps = model.get_attention_maps()
    r = O.forward(sd, x, None, keep_all_maps=True)
    torch.testing.assert_close(r["logits"], y, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(O.get_attention_maps(r["plane_cls"], r["slice_cls"]), ref_maps, rtol=1e-5, atol=1e-10)
    # rollout (dino.py:204-212)
    torch.testing.assert_close(O.get_attention_cls(r["maps"]), ref_rollout, rtol=1e-4, atol=1e-8)
