"""B200-native MST-DINOv2 hot path (DinoV2ClassifierSlice forward + saliency).

Host-side mirror of the reference's `mst.models` surface over a C-ABI CUDA library
(`include/mst_b200.h`, built from `new-vit_b200/csrc/`).  No CPU fallback: every compute entry
point raises if the CUDA library or a GPU is missing.
"""
__all__ = ["DinoV2ClassifierSlice", "ResNetSliceTrans", "SliceTransformerHead", "run_pred", "quantile", "duke_transform", "synth"]


def __getattr__(name):  # lazy: `import new_vit_b200.synth` must not need the CUDA library
    if name in ("DinoV2ClassifierSlice", "run_pred", "quantile", "MSTError"):
        from new_vit_b200 import model as _m
        return getattr(_m, name)
    if name in ("ResNetSliceTrans", "SliceTransformerHead"):
        from new_vit_b200 import resnet_slice as _r
        return getattr(_r, name)
    if name == "duke_transform":
        from new_vit_b200 import transforms as _t
        return _t.duke_transform
    if name == "synth":
        import importlib
        return importlib.import_module("new_vit_b200.synth")
    raise AttributeError(name)
