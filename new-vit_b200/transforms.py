"""Input pipeline in front of `DinoV2ClassifierSlice.forward`, on the GPU.

`duke_transform` is what the reference's `DUKE_Dataset3D` does to a volume between the HDF5 read and the `source` tensor
with its default arguments (mst/data/datasets/dataset_3d_duke.py:36-47: image_resize / resample None, random_center /
flip / random_rotate / noise off): tio.Flip(1) -> CropOrPad(image_crop, padding_mode='minimum')
(mst/data/datasets/augmentations/augmentations_3d.py:144-195) -> ZNormalization(percentiles=(0.5, 99.5),
masking_method=(x > x.min()) & (x < x.max())) (:41-86) -> ImageOrSubjectToTensor (:23-29).  One C-ABI call
(`mst_prepare_volume`, csrc/prep.cu); there is no CPU path.
"""
import torch

from . import _cabi
from ._cabi import MSTError


def duke_transform(data, image_crop=(224, 224, 32), percentiles=(0.5, 99.5), flip=True, check=True, return_stats=False):
    """data: CUDA fp32 tensor [W0,H0,D0], [items,W0,H0,D0] or [items,1,W0,H0,D0] in torchio axis order (W, H, D), all items
    of one shape.  Returns `source` [items, 1, D, H, W] fp32 (image_crop = (W, H, D) as in the reference).

    check=True reads the per-item status back (one small device-to-host copy) and raises RuntimeError where the reference
    does: standard deviation 0 or an empty mask (augmentations_3d.py:75-84).  return_stats=True also returns the float64
    table [items, 8]: min, max, cutoff_lo, cutoff_hi, mean, std, masked voxels, status."""
    if not isinstance(data, torch.Tensor) or data.device.type != "cuda":
        raise MSTError("duke_transform runs on a CUDA tensor only (no CPU fallback)")
    if data.dim() == 3:
        data = data[None]
    if data.dim() == 5:
        if data.shape[1] != 1:
            raise ValueError("duke_transform: one channel per volume (the DUKE 'sub' scan)")
        data = data[:, 0]
    if data.dim() != 4:
        raise ValueError("duke_transform: expected [W,H,D], [items,W,H,D] or [items,1,W,H,D]")
    # raw scanner voxels stay 2 bytes wide until they are on the device (half the host-to-device bytes of the raw-data route)
    raw_code = {torch.int16: 3, torch.uint16: 4}.get(data.dtype, 0)
    data = data.detach().contiguous() if raw_code else data.detach().to(torch.float32).contiguous()
    items, W0, H0, D0 = data.shape
    W, H, D = (int(v) for v in image_crop)
    if (W * H * D) % 4:
        raise ValueError("duke_transform: W*H*D of image_crop must be a multiple of 4")
    L = _cabi.lib()
    ct = _cabi.ctypes
    with torch.cuda.device(data.device):
        out = torch.empty((items, 1, D, H, W), dtype=torch.float32, device=data.device)
        stats = torch.empty((items, 8), dtype=torch.float64, device=data.device)
        need = ct.c_size_t()
        _cabi.check(L.mst_prepare_volume_workspace_bytes(items, W0, H0, D0, ct.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device=data.device)
        stream = ct.c_void_p(torch.cuda.current_stream().cuda_stream)
        _cabi.check(L.mst_prepare_volume(None, _cabi.ptr(data), raw_code, items, W0, H0, D0, W, H, D, 1 if flip else 0,
                                         ct.c_float(percentiles[0] / 100.0), ct.c_float(percentiles[1] / 100.0),
                                         _cabi.ptr(out), _cabi.ptr(stats), _cabi.ptr(ws), need.value, stream))
        if check:
            status = stats[:, 7].cpu()
            if bool((status == 1).any()):
                raise RuntimeError("Standard deviation is 0 for masked values")      # augmentations_3d.py:79-84
            if bool((status == 2).any()):
                raise RuntimeError("quantile() input tensor must be non-empty")      # torch.quantile on an empty selection (:75)
    return (out, stats) if return_stats else out
