"""MST-ResNet (reference mst/models/resnet.py:127-244, `ResNetSliceTrans`) on the shared slice-transformer kernel.

The reference's MST-ResNet is a 2D ResNet per slice (torchvision / MONAI, library code) followed by the SAME slice transformer +
head as MST-DINOv2, with d_model = 512, 16 heads, dim_feedforward = 512 (resnet.py:152-172).  This class keeps that constructor,
`state_dict` layout (`model.*`, `slice_fusion.*`, `cls_token`, `linear.*`), forward and `get_slice_attention` contract; the backbone
stays the library module the reference uses, the slice transformer + head run in `slice_fusion_kernel` (csrc/kernels.cu) through a
head-only handle of the C ABI (`mst_config.depth = 0`, `mst_slice_head_forward`).  Grad-CAM++ of the backbone
(`ResNet.get_attention_maps`, resnet.py:93-122) is backbone-side autograd bookkeeping and is not mirrored."""
import torch
import torch.nn as nn

from . import _cabi
from ._cabi import MSTError
from .model import _SliceFusion


class SliceTransformerHead(nn.Module):
    """cls_token + nn.TransformerEncoder(1 pre-LN layer, ReLU FFN of width emb_ch, final LayerNorm) + Linear on features [B, D, emb_ch]
    (resnet.py:155-172 / dino.py:84-103), evaluated by the CUDA slice-transformer kernel."""

    def __init__(self, emb_ch, nhead, out_ch):
        super().__init__()
        self.emb_ch, self.nhead, self.out_ch = emb_ch, nhead, out_ch
        self.slice_fusion = _SliceFusion(emb_ch, nhead)
        self.cls_token = nn.Parameter(torch.randn(1, 1, emb_ch))
        self.linear = nn.Linear(emb_ch, out_ch)
        self._handle, self._synced = None, None
        self.attention_maps_slice = []

    def _sync(self, dev):
        version = (dev, sum(p._version for p in self.parameters()), tuple(p.data_ptr() for p in self.parameters()))
        if self._handle is not None and self._synced == version:
            return
        if dev.type != "cuda":
            raise MSTError("the slice-transformer head runs on a CUDA device only (no CPU fallback)")
        L = _cabi.lib()
        if self._handle is None:
            cfg = _cabi.MstConfig(self.emb_ch, 0, 0, self.nhead, self.out_ch, 0, _cabi.PRECISION["fp32"], dev.index or 0, 0, 0, 0,
                                  _cabi.FUSION["transformer"], 1, 0, 0, 0.0)
            h = _cabi.ctypes.c_void_p()
            _cabi.check(L.mst_create(_cabi.ctypes.byref(cfg), _cabi.ctypes.byref(h)))
            self._handle = h
        with torch.cuda.device(dev):
            stream = _cabi.ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            keep = []
            for name, t in self.state_dict().items():
                t32 = t.detach().to(device=dev, dtype=torch.float32).contiguous()
                keep.append(t32)
                _cabi.check(L.mst_set_weight(self._handle, name.encode(), _cabi.ptr(t32), t32.numel(), stream))
            _cabi.check(L.mst_finalize_weights(self._handle, stream))
        self._synced = version

    def __del__(self):
        try:
            if self._handle is not None:
                _cabi.lib().mst_destroy(self._handle)
        except Exception:
            pass

    def forward(self, feats, src_key_padding_mask=None, save_attn=False):
        B, D, E = feats.shape
        dev = self.cls_token.device
        self._sync(dev)
        L = _cabi.lib()
        x = feats.detach().to(dev).float().contiguous()
        mask = None
        if src_key_padding_mask is not None:
            mask = src_key_padding_mask.to(dev).to(torch.uint8).contiguous()
        with torch.cuda.device(dev):
            logits = torch.empty((B, self.out_ch), device=dev, dtype=torch.float32)
            slc = torch.empty((B, self.nhead, D + 1), device=dev, dtype=torch.float32) if save_attn else None
            scratch = torch.empty((B, D + 1, E), device=dev, dtype=torch.float32)
            stream = _cabi.ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _cabi.check(L.mst_slice_head_forward(self._handle, _cabi.ptr(x), B, D, _cabi.ptr(mask), _cabi.ptr(logits), None, _cabi.ptr(slc),
                                                 _cabi.ptr(scratch), stream))
        if save_attn:
            self.attention_maps_slice = [slc.unsqueeze(2)]     # row 0 of [B, heads, 1+D, 1+D], what the getter reads (resnet.py:202-203)
        return logits

    def get_slice_attention(self):
        """[B*D, 1, 1]: CLS row without the CLS column, renormalised per head, mean over heads (resnet.py:201-210)."""
        if not self.attention_maps_slice:
            raise IndexError("list index out of range")
        slc = self.attention_maps_slice[-1][:, :, 0, :].contiguous()
        B, heads, Lq = slc.shape
        D = Lq - 1
        with torch.cuda.device(slc.device):
            out = torch.empty((B * D,), device=slc.device, dtype=torch.float32)
            stream = _cabi.ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _cabi.check(_cabi.lib().mst_saliency(self._handle, None, _cabi.ptr(slc), B, D, 1, heads, 1, 1, 1, 14, 14, 0, None, None,
                                                 _cabi.ptr(out), None, None, stream))
        return out[:, None, None]


class ResNetSliceTrans(nn.Module):
    """Drop-in for reference `mst.models.ResNetSliceTrans` (resnet.py:127): per-slice 2D ResNet features -> slice transformer -> logits."""

    def __init__(self, in_ch, out_ch, spatial_dims=2, model=34, pretrained=True, kwargs_resnet={}, rotary_positional_encoding=None,
                 optimizer_kwargs={'lr': 1e-5, 'weight_decay': 1e-2}, backbone=None, **kwargs):
        super().__init__()
        if rotary_positional_encoding is not None:
            raise NotImplementedError("MST-ResNet is built with rotary_positional_encoding=None (the reference's default, resnet.py:135)")
        emb_ch = 512 if model <= 34 else 2048                      # resnet.py:152
        if emb_ch > 1024:
            raise NotImplementedError("the shared slice-transformer kernel holds embeddings up to 1024 wide (ResNet-18/34: 512)")
        if backbone is None:
            if pretrained:
                raise NotImplementedError("pretrained=True downloads torchvision weights (resnet.py:42-44); this machine is offline: "
                                          "pass pretrained=False or a `backbone` module and load a checkpoint")
            import torchvision.models as models                    # the reference's own backbone library (resnet.py:6,14-18)
            backbone = {18: models.resnet18, 34: models.resnet34}[model](weights=None)
            backbone.fc = nn.Identity()                            # emb_ch=None: features, not logits (resnet.py:45-47)
        self.in_ch, self.out_ch, self.spatial_dims, self.optimizer_kwargs = in_ch, out_ch, spatial_dims, optimizer_kwargs
        self.model = backbone
        self.head = SliceTransformerHead(emb_ch, 16, out_ch)       # nhead=16, dim_feedforward=emb_ch (resnet.py:157-160)
        self.attention_maps_slice = []

    # the reference keeps these three directly on the model (resnet.py:155-172); expose the same state_dict keys
    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        return type(sd)((k[len("head."):] if k.startswith("head.") else k, v) for k, v in sd.items())

    def load_state_dict(self, state_dict, strict=True):
        remap = {(k if k.startswith("model.") else "head." + k): v for k, v in state_dict.items()}
        return super().load_state_dict(remap, strict=strict)

    @property
    def device(self):
        return self.head.cls_token.device

    def forward(self, source, src_key_padding_mask=None, **kwargs):
        x = source.to(self.device)                                 # [B, C, D, H, W]   (resnet.py:175)
        B, C, D, H, W = x.shape
        x = x.repeat(1, 3, 1, 1, 1)                                # gray -> RGB        (:182)
        x = x.permute(0, 2, 1, 3, 4).reshape(B * D, 3 * C, H, W)   # '(b d) c h w'      (:183)
        with torch.no_grad():
            feats = self.model(x).reshape(B, D, -1)                # the library backbone (:184-185)
        y = self.head(feats, src_key_padding_mask, save_attn=bool(kwargs.get('save_attn')))
        if kwargs.get('save_attn'):
            self.attention_maps_slice = self.head.attention_maps_slice
        return y

    def get_slice_attention(self):
        return self.head.get_slice_attention()

    def get_attention_maps(self):
        raise NotImplementedError("Grad-CAM++ of the ResNet backbone (resnet.py:93-122, 212-217) is not mirrored; "
                                  "get_slice_attention() is available")
