"""TEST INFRASTRUCTURE ONLY -- runs the reference's OWN input-pipeline classes in this dev container.

`mst/data/datasets/augmentations/augmentations_3d.py` subclasses torchio 0.19.9 (`environment.yaml:25`), which is not in
this image.  The reference's own code in that file -- `CropOrPad._get_six_bounds_parameters` / `apply_transform`
(:166-195), `ZNormalization.apply_normalization` / `_znorm` (:55-86), `ImageOrSubjectToTensor` (:23-29) -- is what
decides the bounds, the percentile cutoffs and the clamp, so it is executed here UNMODIFIED, loaded from the file where
it lies, on top of minimal stand-ins for the torchio base classes it derives from.  The stand-ins restate only what
torchio 0.19.9 does around those overrides:
  * `tio.CropOrPad._compute_center_crop_or_pad`: diff = target - shape; cropping = -min(diff, 0), padding = max(diff, 0),
    each turned into six bounds by `self._get_six_bounds_parameters` (the reference's override),
  * `tio.Pad`: `np.pad(data, ((0,0), (w0,w1), (h0,h1), (d0,d1)), mode=padding_mode)`; `tio.Crop`: slicing + clone,
  * `NormalizationTransform.apply_transform`: mask = masking_method(image.data), then `apply_normalization`,
  * `tio.ZNormalization.znorm` (static): clone().float(), mean / unbiased std of the masked values, None when std == 0,
  * `tio.Flip(axes=1)`: `torch.flip(data, dims=[2])` (the chain's own step, dataset_3d_duke.py:41).
So the fixture this produces (`tests/golden/duke_transform_ref.npz`, made by `tests/golden/make_transform_golden.py`)
pins `oracle/duke_transform_oracle.py` to the reference's own code; the torchio base-class behaviour stays a
restatement ("partially pinned").  `/root/reference` does not exist on the GPU box: nothing that runs there imports this.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("MST_REFERENCE_ROOT", "/root/reference")
_AUG = os.path.join(REFERENCE_ROOT, "mst", "data", "datasets", "augmentations", "augmentations_3d.py")


def reference_available() -> bool:
    return os.path.isfile(_AUG)


class _Image:
    def __init__(self, tensor):
        self.data = tensor
        self.path = None

    def set_data(self, tensor):
        self.data = tensor

    @property
    def shape(self):
        return tuple(self.data.shape)


class _Subject(dict):
    @property
    def spatial_shape(self):
        return tuple(next(iter(self.values())).shape[1:])

    def check_consistent_space(self):
        pass


class _Transform:
    def __init__(self, **kwargs):
        pass

    def __call__(self, subject):
        return self.apply_transform(subject)


class _Pad(_Transform):
    def __init__(self, padding, padding_mode=0, **kwargs):
        self.padding, self.padding_mode = tuple(int(p) for p in padding), padding_mode

    def apply_transform(self, subject):
        p = self.padding
        widths = ((0, 0), p[0:2], p[2:4], p[4:6])
        for image in subject.values():
            if isinstance(self.padding_mode, str):
                out = np.pad(image.data.numpy(), widths, mode=self.padding_mode)
            else:
                out = np.pad(image.data.numpy(), widths, mode="constant", constant_values=self.padding_mode)
            image.set_data(torch.as_tensor(out))
        return subject


class _Crop(_Transform):
    def __init__(self, cropping, **kwargs):
        self.cropping = tuple(int(c) for c in cropping)

    def apply_transform(self, subject):
        i0, i1, j0, j1, k0, k1 = self.cropping
        for image in subject.values():
            _, W, H, D = image.shape
            image.set_data(image.data[:, i0:W - i1, j0:H - j1, k0:D - k1].clone())
        return subject


class _CropOrPad(_Transform):
    def __init__(self, target_shape=None, padding_mode=0, mask_name=None, labels=None, **kwargs):
        self.target_shape, self.padding_mode, self.mask_name, self.labels = target_shape, padding_mode, mask_name, labels
        assert mask_name is None, "the DUKE chain uses the centre rule (no mask)"

    def compute_crop_or_pad(self, subject):
        diff = np.asarray(self.target_shape) - np.asarray(subject.spatial_shape)
        cropping, padding = -np.minimum(diff, 0), np.maximum(diff, 0)
        cropping_params = self._get_six_bounds_parameters(cropping) if cropping.any() else None
        padding_params = self._get_six_bounds_parameters(padding) if padding.any() else None
        return padding_params, cropping_params


class _Normalization(_Transform):
    def __init__(self, masking_method=None, **kwargs):
        self.masking_method = masking_method

    def apply_transform(self, subject):
        for name, image in subject.items():
            mask = self.masking_method(image.data) if callable(self.masking_method) else torch.ones_like(image.data, dtype=torch.bool)
            self.apply_normalization(subject, name, mask)
        return subject


class _ZNormalization(_Normalization):
    @staticmethod
    def znorm(tensor, mask):
        tensor = tensor.clone().float()
        values = tensor[mask]
        mean, std = values.mean(), values.std()
        if std == 0:
            return None
        tensor -= mean
        tensor /= std
        return tensor


class _RescaleIntensity(_Normalization):
    def __init__(self, out_min_max=(0, 1), percentiles=(0, 100), masking_method=None, in_min_max=None, **kwargs):
        super().__init__(masking_method)


def load_reference_augmentations():
    """The reference's augmentations_3d module, executed from its own file over the stand-ins above."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    saved = {k: sys.modules.get(k) for k in ("torchio", "torchio.types", "torchio.transforms", "torchio.transforms.transform", "nibabel")}
    tio = types.ModuleType("torchio")
    tio.Subject, tio.Image, tio.Pad, tio.Crop, tio.CropOrPad = _Subject, _Image, _Pad, _Crop, _CropOrPad
    tio.ZNormalization, tio.RescaleIntensity, tio.EnsureShapeMultiple = _ZNormalization, _RescaleIntensity, _Transform
    tt = types.ModuleType("torchio.types")
    tt.TypeRangeFloat = tt.TypeTripletInt = object
    ttr = types.ModuleType("torchio.transforms")
    ttt = types.ModuleType("torchio.transforms.transform")
    ttt.TypeMaskingMethod = object
    sys.modules.update({"torchio": tio, "torchio.types": tt, "torchio.transforms": ttr, "torchio.transforms.transform": ttt,
                        "nibabel": types.ModuleType("nibabel")})
    try:
        spec = importlib.util.spec_from_file_location("_ref_augmentations_3d", _AUG)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():          # leave no fake torchio behind for other tests
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def reference_duke_transform(volume, image_crop=(224, 224, 32), percentiles=(0.5, 99.5)):
    """dataset_3d_duke.py:36-47 with its default arguments, built from the reference's own CropOrPad / ZNormalization /
    ImageOrSubjectToTensor objects.  volume: numpy fp32 [W0, H0, D0].  Returns `source` [1, D, H, W] torch fp32."""
    A = load_reference_augmentations()
    image = _Image(torch.as_tensor(np.asarray(volume, dtype=np.float32))[None].clone())
    image.set_data(torch.flip(image.data, dims=[2]))                                              # tio.Flip(1)
    subject = _Subject(img=image)
    subject = A.CropOrPad(image_crop, random_center=False, padding_mode="minimum")(subject)       # :42
    subject = A.ZNormalization(per_channel=True, per_slice=False, percentiles=percentiles,
                               masking_method=lambda x: (x > x.min()) & (x < x.max()))(subject)    # :43
    return A.ImageOrSubjectToTensor()(subject["img"]).contiguous()                                # :47 (an Image, not a Subject)
