"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy + torch CPU) of what DUKE_Dataset3D's default transform chain does to one volume between the
HDF5 read and the model's `source` tensor.  Only `tests/` may import it.

PARITY PARTIALLY PINNED.  The chain is built from torchio 0.19.9 classes (pinned in the reference's
`environment.yaml:25`) and torchio is NOT in this image (no network), so the torchio base classes cannot be run here.
The reference's OWN code can: `oracle/ref_transform_harness.py` executes augmentations_3d.py from /root/reference over
minimal stand-ins for those base classes, `tests/golden/duke_transform_ref.npz` holds its outputs (5 seeded volumes; script
`tests/golden/make_transform_golden.py`), and this oracle is bit-equal to them (tests/test_transform_oracle.py).  That pins
the reference's own lines, restated here one by one:
  * `CropOrPad._get_six_bounds_parameters` / `apply_transform` (mst/data/datasets/augmentations/augmentations_3d.py:166-195),
  * `ZNormalization.apply_normalization` / `_znorm` (augmentations_3d.py:55-86: torch.quantile of the masked values, torch.clamp, znorm),
  * `ImageOrSubjectToTensor` (augmentations_3d.py:23-29: swapaxes(1, -1)),
  * the chain and its arguments (mst/data/datasets/dataset_3d_duke.py:36-47).
UNPINNED remainder: what torchio itself does around them, below.
The torchio 0.19.9 parts are restated from its published source and call the SAME library routines torchio calls
(so the arithmetic is the library's, not a re-derivation):
  * `tio.Flip(axes=1)`: `torch.flip(data, dims=[axis + 1])` on the [C, W, H, D] tensor,
  * `tio.Pad(bounds, padding_mode='minimum')`: `np.pad(data, ((0,0), (w0,w1), (h0,h1), (d0,d1)), mode='minimum')`,
  * `tio.Crop(bounds)`: `data[:, i0:W-i1, j0:H-j1, k0:D-k1]`,
  * `CropOrPad._compute_center_crop_or_pad`: padding = max(target - shape, 0), cropping = max(shape - target, 0), pad first,
  * `NormalizationTransform.apply_transform`: the mask comes from `masking_method(tensor)` BEFORE the clamp,
  * `ZNormalization.znorm`: `values = tensor.clone().float()[mask]; mean, std = values.mean(), values.std()` (unbiased);
    `std == 0` -> None (the reference then raises RuntimeError); `tensor -= mean; tensor /= std`.
"""
import numpy as np
import torch


def six_bounds(numbers):
    """augmentations_3d.py:166-175 with random_center=False: ini = ceil(n / 2), fin = n - ini."""
    out = []
    for n in numbers:
        ini = int(np.ceil(n / 2))
        out.extend([ini, int(n) - ini])
    return tuple(out)


def crop_or_pad(data, target, padding_mode="minimum"):
    """data [C, W, H, D] numpy -> [C, *target] (augmentations_3d.py:178-195 + tio.CropOrPad._compute_center_crop_or_pad)."""
    shape = np.array(data.shape[1:])
    diff = np.array(target) - shape
    cropping = -np.minimum(diff, 0)
    padding = np.maximum(diff, 0)
    if padding.any():
        w0, w1, h0, h1, d0, d1 = six_bounds(padding)
        data = np.pad(data, ((0, 0), (w0, w1), (h0, h1), (d0, d1)), mode=padding_mode)      # tio.Pad
    if cropping.any():
        i0, i1, j0, j1, k0, k1 = six_bounds(cropping)
        W, H, D = data.shape[1:]
        data = data[:, i0:W - i1, j0:H - j1, k0:D - k1]                                      # tio.Crop
    return np.ascontiguousarray(data)


def znorm_percentile(image, percentiles=(0.5, 99.5)):
    """ZNormalization(per_channel=True, per_slice=False, masking_method=lambda x: (x > x.min()) & (x < x.max()),
    percentiles=...) on one channel (augmentations_3d.py:55-86; dataset_3d_duke.py:43).  image: torch fp32 [1, W, H, D].
    Returns (normalised tensor, dict of the intermediate statistics)."""
    mask = (image > image.min()) & (image < image.max())
    image_data = image.clone()
    cutoff = torch.quantile(image_data.masked_select(mask).float(), torch.tensor(percentiles) / 100.0)   # :75
    lo, hi = cutoff.to(image_data.dtype).tolist()
    torch.clamp(image_data, lo, hi, out=image_data)                                                      # :76
    tensor = image_data.clone().float()                                                                  # tio znorm
    values = tensor[mask]
    mean, std = values.mean(), values.std()
    if std == 0:
        raise RuntimeError("Standard deviation is 0 for masked values")                                  # :79-84
    tensor -= mean
    tensor /= std
    stats = dict(min=float(image.min()), max=float(image.max()), lo=lo, hi=hi, mean=float(mean), std=float(std),
                 count=int(mask.sum()))
    return tensor, stats


def duke_transform(volume, image_crop=(224, 224, 32), percentiles=(0.5, 99.5)):
    """dataset_3d_duke.py:36-47 (image_resize / resample None, no random augmentation) on one volume.
    volume: numpy fp32 [W0, H0, D0] (the HDF5 array, torchio's [W, H, D]).  Returns (source [1, D, H, W] torch fp32, stats)."""
    data = torch.as_tensor(np.asarray(volume, dtype=np.float32))[None]            # tio.ScalarImage(tensor=data[None]) [C,W,H,D]
    data = torch.flip(data, dims=[2])                                             # tio.Flip(1)
    data = torch.as_tensor(crop_or_pad(data.numpy(), image_crop, "minimum"))      # CropOrPad(..., padding_mode='minimum')
    data, stats = znorm_percentile(data, percentiles)                             # ZNormalization
    return data.swapaxes(1, -1).contiguous(), stats                               # ImageOrSubjectToTensor
