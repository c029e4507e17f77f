"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the host mirror keeps the
reference's state_dict layout and error behaviour, synthetic data is deterministic."""
import ctypes

import pytest
import torch

from new_vit_b200 import _cabi, synth


def test_library_loads_and_exports_all_declared_symbols():
    L = _cabi.lib()
    names = _cabi.declared_symbols()
    assert len(names) >= 14 and "mst_forward" in names and "mst_saliency" in names
    for n in names:
        assert hasattr(L, n), n
    assert L.mst_abi_version() == _cabi.ABI_VERSION == 4


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must fail loudly."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _cabi.lib()
    cfg = _cabi.MstConfig(384, 12, 6, 12, 2, 257, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0.1)
    h = ctypes.c_void_p()
    assert L.mst_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"no CUDA device" in L.mst_last_error()
    from new_vit_b200 import DinoV2ClassifierSlice
    from new_vit_b200._cabi import MSTError
    m = DinoV2ClassifierSlice(1, 2, pretrained=False)
    with pytest.raises(MSTError):
        m(torch.zeros(1, 1, 2, 224, 224))


def test_state_dict_layout_matches_reference():
    from new_vit_b200 import DinoV2ClassifierSlice
    m = DinoV2ClassifierSlice(in_ch=1, out_ch=2, pretrained=False)
    sd = m.state_dict()
    ref = synth.make_state_dict("s", 2, seed=0)
    assert list(sd.keys()) == list(ref.keys()) and len(sd) == 168
    for k in sd:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    assert m.encoder.num_features == 384 and m.emb_ch == 384
    m.load_state_dict(ref)
    assert torch.equal(m.state_dict()["encoder.blocks.0.3.attn.qkv.weight"], ref["encoder.blocks.0.3.attn.qkv.weight"])
    from oracle import ref_harness
    if ref_harness.reference_available():
        real = ref_harness.build_reference_model()
        assert list(real.state_dict().keys()) == list(sd.keys())
        real.load_state_dict(sd)  # our checkpoints load into the reference and vice versa


def test_unbuilt_options_raise():
    from new_vit_b200 import DinoV2ClassifierSlice
    for kw in (dict(pretrained=True),):
        args = dict(in_ch=1, out_ch=2, pretrained=False)
        args.update(kw)
        with pytest.raises(NotImplementedError):
            DinoV2ClassifierSlice(**args)
    with pytest.raises(ValueError):
        DinoV2ClassifierSlice(1, 2, pretrained=False, slice_fusion="max")
    with pytest.raises(ValueError):   # transformer_blocks.py:358
        DinoV2ClassifierSlice(1, 2, pretrained=False, rotary_positional_encoding="xpos")


@pytest.mark.parametrize("kw", [dict(use_bottleneck=True), dict(use_slice_pos_emb=True), dict(slice_fusion="linear"),
                                dict(slice_fusion="average", enable_linear=False), dict(use_bottleneck=True, slice_fusion="linear"),
                                dict(rotary_positional_encoding="RoPE", use_bottleneck=True)])
def test_constructor_variants_keep_the_reference_state_dict_layout(kw):
    """dino.py:75-103: bottleneck, slice position embedding, slice_fusion, enable_linear add / drop / resize tensors."""
    from new_vit_b200 import DinoV2ClassifierSlice
    from oracle import ref_harness
    m = DinoV2ClassifierSlice(in_ch=1, out_ch=3, pretrained=False, **kw)
    sd = m.state_dict()
    skw = {k: v for k, v in kw.items() if k != "rotary_positional_encoding"}
    syn = synth.make_state_dict("s", 3, seed=1, rope=kw.get("rotary_positional_encoding") == "RoPE", **skw)
    assert sorted(sd.keys()) == sorted(syn.keys())
    assert all(tuple(sd[k].shape) == tuple(syn[k].shape) for k in sd)
    if ref_harness.reference_available():
        real = ref_harness.build_reference_model(out_ch=3, **kw)
        rsd = real.state_dict()
        assert sorted(rsd.keys()) == sorted(sd.keys())
        assert all(tuple(rsd[k].shape) == tuple(sd[k].shape) for k in sd)
        assert real.emb_ch == m.emb_ch


def test_register_architecture_layout():
    from new_vit_b200 import DinoV2ClassifierSlice
    m = DinoV2ClassifierSlice(1, 2, pretrained=False, use_registers=True, hub_layout=True, img_size=518)
    sd = m.state_dict()
    assert tuple(sd["encoder.register_tokens"].shape) == (1, 4, 384)
    assert tuple(sd["encoder.pos_embed"].shape) == (1, 1370, 384)      # hub checkpoints: 37 x 37 grid @518
    assert "encoder.blocks.11.ls2.gamma" in sd and "encoder.blocks.0.0.norm1.weight" not in sd


def test_synth_is_deterministic():
    a, b = synth.make_state_dict("s", 2, seed=4, variant="peaky"), synth.make_state_dict("s", 2, seed=4, variant="peaky")
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(synth.make_volume(1, 3, 28, 28, seed=9), synth.make_volume(1, 3, 28, 28, seed=9))
    m = synth.make_padding_mask(4, 32, seed=1)
    assert m.shape == (4, 32) and not m[0].any() and m[1:].any()


def test_checkpoint_loading_api_mirrors_base_model(tmp_path):
    """base_model.py:50-81 + LightningModule.load_from_checkpoint, as main_predict.py:215 uses them: a Lightning-style
    checkpoint ('hyper_parameters' + 'state_dict') and best_checkpoint.json."""
    from new_vit_b200 import DinoV2ClassifierSlice
    sd = synth.make_state_dict("s", 2, seed=3)
    hp = dict(in_ch=1, out_ch=2, spatial_dims=2, pretrained=False, model_size="s", slice_fusion="transformer",
              loss_kwargs={}, aucroc_kwargs={"task": "binary"})                    # BasicClassifier's extra arguments are swallowed
    torch.save({"state_dict": sd, "hyper_parameters": hp, "epoch": 7}, tmp_path / "epoch=7.ckpt")
    DinoV2ClassifierSlice.save_best_checkpoint(tmp_path, tmp_path / "epoch=7.ckpt")
    assert DinoV2ClassifierSlice._get_best_checkpoint_path(tmp_path) == tmp_path / "epoch=7.ckpt"
    m = DinoV2ClassifierSlice.load_best_checkpoint(tmp_path)
    got = m.state_dict()
    assert list(got.keys()) == list(sd.keys()) and all(torch.equal(got[k], sd[k]) for k in sd)
    assert m._dirty                                                                # the next forward re-packs the weights

    # load_pretrained on a directory, and load_weights with a filter (only the encoder is taken)
    fresh = DinoV2ClassifierSlice(1, 2, pretrained=False)
    before = {k: v.clone() for k, v in fresh.state_dict().items()}
    fresh.load_weights(sd, filter=lambda key: key.startswith("encoder."))
    after = fresh.state_dict()
    assert torch.equal(after["encoder.pos_embed"], sd["encoder.pos_embed"])
    assert torch.equal(after["linear.weight"], before["linear.weight"])
    assert fresh.load_pretrained(tmp_path) is fresh
    assert torch.equal(fresh.state_dict()["linear.weight"], sd["linear.weight"])


def test_checkpoint_of_a_hub_trained_model_needs_no_download(tmp_path):
    """A checkpoint trained with pretrained=True (hub encoder: LayerScale, blocks.<i>, 518-pixel position table, registers)
    carries every weight; the architecture is read off its tensors."""
    from new_vit_b200 import DinoV2ClassifierSlice
    sd = synth.make_state_dict("s", 2, seed=4, img_size=518, layerscale=True, chunked_names=False, num_registers=4)
    torch.save({"state_dict": sd, "hyper_parameters": dict(in_ch=1, out_ch=2, pretrained=True, use_registers=True)},
               tmp_path / "hub.ckpt")
    m = DinoV2ClassifierSlice.load_from_checkpoint(tmp_path / "hub.ckpt", precision="fp32")
    assert m.num_registers == 4 and m.encoder.pos_embed.shape[1] == 1370 and m.precision == "fp32"
    assert "encoder.blocks.11.ls2.gamma" in m.state_dict()
    assert all(torch.equal(m.state_dict()[k], sd[k]) for k in sd)
