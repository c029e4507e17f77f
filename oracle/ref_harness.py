"""TEST INFRASTRUCTURE ONLY -- loads the *real* reference model in this dev container.

The reference (`/root/reference`, read-only) is a Python package whose import chain pulls in
packages that are not installed here (`pytorch_lightning`, `torchmetrics`, `monai`,
`matplotlib`).  None of them is on the hot path, so they are replaced by inert stub modules
before `mst.models.dino` is imported (recipe: SURVEY.md section 8c).

This module is used by `tests/golden/make_golden.py` to generate the committed golden vectors
and by the (container-only) oracle-vs-reference pin test.  `/root/reference` does not exist on
the GPU box: nothing that runs there may import this file.
"""
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("MST_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mst", "models"))


def _install_stubs() -> None:
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

            @property
            def device(self):
                return next(self.parameters()).device

        class LightningDataModule:
            pass

        pl.LightningModule = LightningModule
        pl.LightningDataModule = LightningDataModule
        sys.modules["pytorch_lightning"] = pl
        ut = types.ModuleType("pytorch_lightning.utilities")
        sys.modules["pytorch_lightning.utilities"] = ut

    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")

        class _Metric(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

            def update(self, *a, **k):
                pass

            def compute(self):
                return torch.tensor(0.0)

            def reset(self):
                pass

        tm.AUROC = tm.Accuracy = tm.MeanSquaredError = tm.MeanAbsoluteError = _Metric
        sys.modules["torchmetrics"] = tm

    if "monai" not in sys.modules:
        monai = types.ModuleType("monai")
        networks = types.ModuleType("monai.networks")
        nets = types.ModuleType("monai.networks.nets")
        for n in ("resnet10", "resnet18", "resnet34", "resnet50", "resnet101", "resnet152",
                  "ResNetFeatures", "ResNet", "ResNetBlock", "ResNetBottleneck"):
            setattr(nets, n, None)
        monai.networks = networks
        networks.nets = nets
        sys.modules["monai"] = monai
        sys.modules["monai.networks"] = networks
        sys.modules["monai.networks.nets"] = nets

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.get_cmap = lambda *a, **k: None
        mpl.pyplot = plt
        mpl.use = lambda *a, **k: None
        mpl.colormaps = {}
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def load_reference_class():
    """Return the reference's own `DinoV2ClassifierSlice` class (mst/models/dino.py:32)."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from mst.models.dino import DinoV2ClassifierSlice  # noqa: E402

    return DinoV2ClassifierSlice


def build_reference_model(state_dict=None, out_ch=2, model_size="s", hub_layout=False, num_registers=0,
                          pos_img_size=None, **kw):
    """`hub_layout=True` swaps in the encoder configuration of the torch.hub DINOv2 checkpoints that
    `pretrained=True` would download (dino.py:59-63): LayerScale (init_values=1.0) and un-chunked block names
    (block_chunks=0), built from the reference's own vendored factory (vision_transformer.py:340-352);
    `num_registers=4` is the "_reg" hub architecture (dino.py:60-61) and `pos_img_size=518` the hub checkpoints'
    position-embedding size.  Other keyword arguments go to the reference constructor unchanged
    (use_bottleneck, use_slice_pos_emb, slice_fusion, enable_linear)."""
    Cls = load_reference_class()
    model = Cls(in_ch=1, out_ch=out_ch, pretrained=False, model_size=model_size, use_registers=num_registers > 0, **kw).eval()
    if hub_layout or num_registers or pos_img_size:
        from mst.models.extern.dinov2 import vision_transformer as vits
        factory = {"s": vits.vit_small, "b": vits.vit_base, "l": vits.vit_large}[model_size]
        extra = dict(init_values=1.0, block_chunks=0) if hub_layout else {}
        if num_registers:   # torch.hub's dinov2_vit*14_reg entry points (what dino.py:60-61 loads) pass these two
            extra.update(interpolate_antialias=True, interpolate_offset=0.0)
        model.encoder = factory(patch_size=14, img_size=pos_img_size or 224, num_register_tokens=num_registers, **extra).eval()
    if state_dict is not None:
        pe = state_dict["encoder.pos_embed"]
        if tuple(pe.shape) != tuple(model.encoder.pos_embed.shape):
            # the local encoder is built for img_size 224 (vision_transformer.py:340-352); a
            # pos_embed sized for another square input makes interpolate_pos_encoding a no-op
            model.encoder.pos_embed = nn.Parameter(torch.zeros_like(pe))
        missing, unexpected = model.load_state_dict(state_dict, strict=True)
        assert not missing and not unexpected
    return model
