"""Host-side mirror of the reference's `mst.models.DinoV2ClassifierSlice` (reference
mst/models/dino.py:32-275) and of `run_pred` (scripts/main_predict.py:55-164) over the C-ABI CUDA
library.  Same constructor arguments, same `state_dict` key layout, same forward / getter
signatures and error behaviour; the arithmetic runs in libmst_b200.so (hand-written sm_100a
kernels).  PyTorch is used for device memory, streams and parameter bookkeeping only.

Constructor variants of dino.py:56-103 are mirrored: use_registers (hub "_reg" architecture, 4 register
tokens), use_bottleneck, use_slice_pos_emb, slice_fusion in {'transformer','linear','average'},
enable_linear; `img_size` is the input size `encoder.pos_embed` is built for (224 local factory, 518 hub
checkpoints) -- other input sizes get the bicubically resampled table (vision_transformer.py:179-211).
rotary_positional_encoding='RoPE' rotates the slice-token queries and keys (transformer_blocks.py:262-264,335-351).
Not mirrored (raise NotImplementedError): pretrained=True (downloads hub weights; offline) and
rotary_positional_encoding='LiRE' (transformer_blocks.py:352-358; see DESIGN.md for what the reference's version does).
"""
import json
from pathlib import Path

import torch
import torch.nn as nn

from new_vit_b200 import _cabi, synth
from new_vit_b200._cabi import MSTError  # noqa: F401
from new_vit_b200.training import SLICE_PARAM_NAMES, EncoderFunction, LightningSurface, SliceHeadFunction


# ---- parameter containers that reproduce the reference's module tree (names only; never called) ----
class _Attn(nn.Module):
    def __init__(self, E):
        super().__init__()
        self.qkv = nn.Linear(E, 3 * E)
        self.proj = nn.Linear(E, E)


class _Mlp(nn.Module):
    def __init__(self, E):
        super().__init__()
        self.fc1 = nn.Linear(E, 4 * E)
        self.fc2 = nn.Linear(4 * E, E)


class _LayerScale(nn.Module):
    def __init__(self, E):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(E))


class _Block(nn.Module):
    def __init__(self, E, layerscale=False):
        super().__init__()
        self.norm1 = nn.LayerNorm(E, eps=1e-6)
        self.attn = _Attn(E)
        if layerscale:
            self.ls1 = _LayerScale(E)
        self.norm2 = nn.LayerNorm(E, eps=1e-6)
        self.mlp = _Mlp(E)
        if layerscale:
            self.ls2 = _LayerScale(E)


class _PatchEmbed(nn.Module):
    def __init__(self, E):
        super().__init__()
        self.proj = nn.Conv2d(3, E, kernel_size=14, stride=14)


class _Encoder(nn.Module):
    """Key layout of DinoVisionTransformer: local factory (block_chunks=1 => blocks.0.<i>, no LayerScale) or, with
    hub_layout, the torch.hub checkpoints' (blocks.<i>, ls1/ls2.gamma)."""

    def __init__(self, E, depth, heads, pos_tokens, hub_layout=False, num_registers=0):
        super().__init__()
        self.num_features = self.embed_dim = E
        self.num_heads = heads
        self.num_register_tokens = num_registers
        self.cls_token = nn.Parameter(torch.zeros(1, 1, E))
        self.pos_embed = nn.Parameter(torch.zeros(1, pos_tokens, E))
        if num_registers:
            self.register_tokens = nn.Parameter(torch.zeros(1, num_registers, E))
        self.mask_token = nn.Parameter(torch.zeros(1, E))
        self.patch_embed = _PatchEmbed(E)
        self.depth = depth
        if hub_layout:
            self.blocks = nn.ModuleList([_Block(E, layerscale=True) for _ in range(depth)])
        else:
            self.blocks = nn.ModuleList([nn.ModuleList([_Block(E) for _ in range(depth)])])
        self.norm = nn.LayerNorm(E, eps=1e-6)


class _Rotary(nn.Module):
    # RotaryEmbedding(dim=head_dim, theta=256, freqs_for='lang') keeps its frequencies as a (frozen) parameter
    # (rotary_embedding_torch.py:104,117; transformer_blocks.py:335-351)
    def __init__(self, hd):
        super().__init__()
        self.freqs = nn.Parameter(torch.zeros(hd // 2), requires_grad=False)


class _Liere(nn.Module):
    # AttentionLiereRotator(head_dim, liere_block_size=head_dim//2, spacial_dims=1, axes_length=33, num_heads)
    # (transformer_blocks.py:352-358; rotary_embedding_torch.py:329-344): head_dim / block = 2 trainable generator tables
    def __init__(self, hd):
        super().__init__()
        blk = hd // 2
        self.vars = nn.ParameterList([nn.Parameter(torch.randn((blk * blk - blk) // 2, 33, 1)) for _ in range(hd // blk)])


class _SliceLayer(nn.Module):
    def __init__(self, E, heads, rope=False, liere=False):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(E, heads, dropout=0.0, batch_first=True)
        if rope:
            self.self_attn.rotary_positional_encoding = _Rotary(E // heads)
        if liere:
            self.self_attn.rotary_positional_encoding = _Liere(E // heads)
        self.linear1 = nn.Linear(E, E)
        self.linear2 = nn.Linear(E, E)
        self.norm1 = nn.LayerNorm(E)
        self.norm2 = nn.LayerNorm(E)


class _SliceFusion(nn.Module):
    def __init__(self, E, heads, rope=False, liere=False):
        super().__init__()
        self.layers = nn.ModuleList([_SliceLayer(E, heads, rope, liere)])
        self.norm = nn.LayerNorm(E)


class DinoV2ClassifierSlice(LightningSurface, nn.Module):
    """Drop-in for reference `mst.models.DinoV2ClassifierSlice` (dino.py:32), including the Lightning-module methods of its
    base classes that the training script drives (base_model.py; new_vit_b200/training.py)."""

    def __init__(self, in_ch, out_ch, spatial_dims=2, pretrained=True, save_attn=False,
                 rotary_positional_encoding=None, optimizer_kwargs={'lr': 1e-6, 'weight_decay': 1e-2},
                 model_size='s', use_registers=False, use_bottleneck=False, use_slice_pos_emb=False,
                 enable_linear=True, enable_trans=True, slice_fusion='transformer', freeze=False,
                 precision='bf16', img_size=224, hub_layout=False, **kwargs):
        super().__init__()
        if pretrained:
            raise NotImplementedError(
                "pretrained=True downloads hub weights (dino.py:59-63); this machine is offline. "
                "Construct with pretrained=False and load a checkpoint with load_state_dict().")
        if rotary_positional_encoding not in (None, 'RoPE', 'LiRE'):
            raise ValueError(f"Unkown parameter {rotary_positional_encoding} for rotary_positional_encoding")  # transformer_blocks.py:358
        if rotary_positional_encoding in ('RoPE', 'LiRE') and slice_fusion != 'transformer':
            rotary_positional_encoding = None               # only the transformer fusion has attention (dino.py:84-96)
        if slice_fusion not in _cabi.FUSION:
            raise ValueError(f"slice_fusion {slice_fusion!r} unsupported")
        if model_size not in synth.VIT_CFG:
            raise ValueError(f"model_size {model_size!r} unsupported")
        if precision not in _cabi.PRECISION:
            raise ValueError("precision must be 'bf16' or 'fp32'")
        E, depth, heads = synth.VIT_CFG[model_size]
        self.in_ch, self.out_ch, self.spatial_dims = in_ch, out_ch, spatial_dims
        self.save_attn = save_attn
        self.attention_maps = []
        self.attention_maps_slice = []
        self.use_registers = use_registers
        self.slice_fusion_type = slice_fusion
        self.rotary = rotary_positional_encoding
        self.precision = precision
        self.model_size = model_size
        pos_tokens = 1 + (img_size // 14) ** 2
        self.num_registers = 4 if use_registers else 0     # dinov2_vit*14_reg (dino.py:60-61)
        self.encoder = _Encoder(E, depth, heads, pos_tokens, hub_layout=hub_layout, num_registers=self.num_registers)
        emb = E
        if use_bottleneck:                                  # dino.py:75-77
            self.bottleneck = nn.Linear(E, E // 4)
            emb = E // 4
        self.emb_ch = emb
        self.enable_linear = bool(enable_linear)
        if slice_fusion == 'transformer':                   # dino.py:80-97
            if use_slice_pos_emb:
                self.slice_pos_emb = nn.Embedding(256, emb)
            self.slice_fusion = _SliceFusion(emb, synth.SLICE_HEADS, rope=rotary_positional_encoding == 'RoPE',
                                             liere=rotary_positional_encoding == 'LiRE')
            self.cls_token = nn.Parameter(torch.zeros(1, 1, emb))
        head_in = emb * 32 if slice_fusion == 'linear' else emb   # dino.py:98-99
        self.linear = nn.Linear(head_in, out_ch) if enable_linear else nn.Identity()
        # the reference's init (zero encoder biases, unit LayerNorm, default-initialised conv / slice layers), seeded
        # from the global torch RNG
        sd = synth.make_state_dict(model_size, out_ch, seed=int(torch.randint(0, 2 ** 31 - 1, (1,)).item()),
                                   img_size=img_size, layerscale=hub_layout, chunked_names=not hub_layout,
                                   num_registers=self.num_registers, use_bottleneck=use_bottleneck,
                                   use_slice_pos_emb=use_slice_pos_emb and slice_fusion == 'transformer',
                                   slice_fusion=slice_fusion, enable_linear=enable_linear,
                                   rope=rotary_positional_encoding == 'RoPE', strict_init=True)
        if rotary_positional_encoding == 'LiRE':            # its generator tables keep their own torch.randn init (:343)
            sd.update({k: v.detach().clone() for k, v in self.state_dict().items() if ".rotary_positional_encoding.vars." in k})
        nn.Module.load_state_dict(self, sd, strict=True)
        self._init_lightning_surface(out_ch, optimizer=kwargs.get("optimizer"), optimizer_kwargs=optimizer_kwargs,
                                     lr_scheduler=kwargs.get("lr_scheduler"), lr_scheduler_kwargs=kwargs.get("lr_scheduler_kwargs"),
                                     loss=kwargs.get("loss", nn.CrossEntropyLoss), loss_kwargs=kwargs.get("loss_kwargs"))
        # DinoVisionTransformer(interpolate_antialias, interpolate_offset): the vendored factory and the plain hub checkpoints
        # resample with (False, 0.1) (vision_transformer.py:66-67); the hub "_reg" models use_registers loads (dino.py:60-61) are
        # built with (True, 0.0).  Constructor keywords override, as they do on the reference's own factory.
        self.interpolate_antialias = bool(kwargs.get("interpolate_antialias", use_registers))
        self.interpolate_offset = float(kwargs.get("interpolate_offset", 0.0 if use_registers else 0.1))
        if freeze:
            for p in self.encoder.parameters():
                p.requires_grad = False
        self._handle = None
        self._handle_key = None
        self._dirty = True
        self._workspace = None
        self._h2d = None
        self._copy_stream = None
        # Host batches are moved in chunks of whole volumes, H2D of chunk k+1 overlapped with the kernels of chunk k.
        # Only the first chunk's copy is exposed, so it is small; later chunks grow so that the GEMM grids keep full
        # waves (8 volumes = 514 m-tiles = 3.5 waves of 148 SMs cost 13 % in wave quantisation; 32 volumes cost 1.5 %).
        # (profiles/e2e_schedule.py, 64 volumes: (4,12,48) 31.6 ms, (4,12,16,32) 32.6, (8,24,32) 31.9, device-resident 28.7)
        self.h2d_chunk_volumes = (4, 12, 48)   # then the last size repeats; an int = fixed chunk size; 0/None = off
        # Forwards of at most this many slices (B*D, x8 with TTA) are launch-bound (one volume = 77 kernels of ~10 us): they run
        # from persistent input / output buffers so that the C library can replay them as one CUDA graph
        # (mst_set_graph_threshold); results are returned as copies.  0 = always launch eagerly.
        self.graph_max_slices = 256
        self._static = {}
        self._last = None
        self._last_inputs = None
        self.register_load_state_dict_post_hook(lambda m, k: setattr(m, "_dirty", True))

    # -- Lightning-style conveniences the callers use (base_model.py) --------------------------------
    @property
    def device(self):
        return self.encoder.cls_token.device

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._dirty = True
        return r

    # -- checkpoints (base_model.py:50-81; main_predict.py:215 calls load_best_checkpoint) -------------
    _BEST = 'best_checkpoint.json'          # {'best_model_epoch': <file name>} next to the checkpoints (base_model.py:51-60)

    @classmethod
    def save_best_checkpoint(cls, path_checkpoint_dir, best_model_path):
        (Path(path_checkpoint_dir) / cls._BEST).write_text(json.dumps({'best_model_epoch': Path(best_model_path).name}))

    @classmethod
    def _get_best_checkpoint_path(cls, path_checkpoint_dir, **kwargs):
        run = Path(path_checkpoint_dir)
        return run / json.loads((run / cls._BEST).read_text())['best_model_epoch']

    @classmethod
    def load_best_checkpoint(cls, path_checkpoint_dir, **kwargs):
        """base_model.py:62-65: the checkpoint best_checkpoint.json names, through load_from_checkpoint."""
        return cls.load_from_checkpoint(cls._get_best_checkpoint_path(path_checkpoint_dir), **kwargs)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=True, **kwargs):
        """What `LightningModule.load_from_checkpoint` does for the reference class: constructor arguments from the
        checkpoint's 'hyper_parameters' (written by `save_hyperparameters()`, base_model.py:13-14), overridden by **kwargs,
        then its 'state_dict'.  A checkpoint trained with `pretrained=True` was built on the torch.hub encoder
        (dino.py:59-63: LayerScale, `blocks.<i>` names, a 518-pixel position table, optional registers); nothing is
        downloaded here -- that architecture is read off the checkpoint's own tensors and every weight comes from it."""
        checkpoint = torch.load(checkpoint_path, map_location=map_location or 'cpu', weights_only=False)
        state = checkpoint['state_dict']
        hparams = dict(checkpoint.get('hyper_parameters', {}))
        hparams.update(kwargs)
        hparams['pretrained'] = False
        pos = state['encoder.pos_embed']
        hparams.setdefault('img_size', 14 * int(round((pos.shape[1] - 1) ** 0.5)))
        hparams.setdefault('hub_layout', 'encoder.blocks.0.ls1.gamma' in state)
        hparams['use_registers'] = 'encoder.register_tokens' in state
        size = {v[0]: k for k, v in synth.VIT_CFG.items()}.get(pos.shape[-1])
        if size is not None:
            hparams['model_size'] = size
        model = cls(**hparams)
        model.load_state_dict(state, strict=strict)
        return model

    def load_pretrained(self, checkpoint_path, map_location=None, **kwargs):
        """base_model.py:67-73: a checkpoint file, or a run directory (then its best checkpoint), into THIS model."""
        path = Path(checkpoint_path)
        if path.is_dir():
            path = self._get_best_checkpoint_path(path, **kwargs)
        return self.load_weights(torch.load(path, map_location=map_location, weights_only=False)["state_dict"], **kwargs)

    def load_weights(self, pretrained_weights, strict=True, **kwargs):
        """base_model.py:75-81: overlay the tensors `filter(key)` keeps (default: all of them) on the current state_dict
        and load the result, so keys the filter drops keep their present values."""
        keep = kwargs.get('filter') or (lambda key: True)
        merged = self.state_dict()
        merged.update({k: v for k, v in pretrained_weights.items() if keep(k)})
        self.load_state_dict(merged, strict=strict)
        return self

    # -- weights -> C handle --------------------------------------------------------------------------
    def _release(self):
        if self._handle is not None:
            _cabi.lib().mst_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def sync_weights(self):
        """(Re)pack the current parameters into the CUDA library (call after in-place edits)."""
        dev = self.device
        if dev.type != "cuda":
            raise MSTError("DinoV2ClassifierSlice runs on a CUDA device only (no CPU fallback): call .cuda() first")
        L = _cabi.lib()
        E = self.encoder.embed_dim
        key = (dev.index or 0, self.precision, self.encoder.pos_embed.shape[1])
        if self._handle is None or self._handle_key != key:
            self._release()
            cfg = _cabi.MstConfig(E, self.encoder.depth, self.encoder.num_heads, synth.SLICE_HEADS, self.out_ch,
                                  self.encoder.pos_embed.shape[1], _cabi.PRECISION[self.precision], key[0],
                                  self.num_registers, int(hasattr(self, "bottleneck")), int(hasattr(self, "slice_pos_emb")),
                                  _cabi.FUSION[self.slice_fusion_type], int(self.enable_linear),
                                  {None: 0, 'RoPE': 1, 'LiRE': 2}[self.rotary],
                                  int(self.interpolate_antialias), self.interpolate_offset)
            h = _cabi.ctypes.c_void_p()
            _cabi.check(L.mst_create(_cabi.ctypes.byref(cfg), _cabi.ctypes.byref(h)))
            self._handle, self._handle_key = h, key
            self._static = {}
        with torch.cuda.device(dev):
            stream = _cabi.ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            names, keep = [], []
            for name, t in self.state_dict().items():
                names.append(name.encode())
                keep.append(t.detach().to(device=dev, dtype=torch.float32).contiguous())
            n = len(names)
            _cabi.check(L.mst_set_weights(self._handle, n, (_cabi.ctypes.c_char_p * n)(*names),
                                          (_cabi.ctypes.c_void_p * n)(*[t.data_ptr() for t in keep]),
                                          (_cabi.ctypes.c_int64 * n)(*[t.numel() for t in keep]), stream))
            _cabi.check(L.mst_finalize_weights(self._handle, stream))
        self._dirty = False
        self._synced_version = (self._param_version(encoder_only=True), self._param_version())
        self._params_stepped = False

    # -- forward (dino.py:110-167) ---------------------------------------------------------------------
    def _param_version(self, encoder_only=False):
        return sum(p._version for p in (self.encoder.parameters() if encoder_only else self.parameters()))

    def forward(self, source, save_attn=False, src_key_padding_mask=None, **kwargs):
        if source.dim() != 5:
            raise ValueError(f"expected source [B, C, D, H, W], got {tuple(source.shape)}")
        if self.rotary == 'LiRE':
            # rotary_positional_encoding='LiRE' in the reference: AttentionLiereRotator hard-codes 33 tokens
            # (rotary_embedding_torch.py:350) and returns a permuted [B, L, heads, hd] tensor that the caller can .view() as
            # [B*heads, L, hd] for batch 1 only (transformer_blocks.py:263).  Everything else raises there, and here, with the same
            # messages (tests/test_training_cpu.py pins them against the live reference).
            B_, L_, hd_ = source.shape[0], source.shape[2] + 1, self.emb_ch // synth.SLICE_HEADS
            if L_ != 33:
                raise RuntimeError(f"shape '[{B_}, 33, {synth.SLICE_HEADS}, {hd_}]' is invalid for input of size "
                                   f"{B_ * synth.SLICE_HEADS * L_ * hd_}")
            if B_ != 1:
                raise RuntimeError("view size is not compatible with input tensor's size and stride (at least one dimension spans "
                                   "across two contiguous subspaces). Use .reshape(...) instead.")
            # batch 1, 32 slices: the one case the reference evaluates (csrc/kernels.cu, LiRE branch of slice_fusion_kernel)
        # In-place parameter updates (an optimizer step) do not pass through load_state_dict: they are detected by tensor version
        # (torch optimizers) or by the flag FusedAdamW raises.  The training path reads the slice transformer's live parameters
        # and needs only the (frozen) encoder packed, so optimizer steps do not trigger a re-pack there.
        train_path = self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        encoder_only = (train_path or bool(kwargs.get("_encoder_only", False))) and not any(p.requires_grad for p in self.encoder.parameters())
        stale = (self._param_version(encoder_only=True) != getattr(self, "_synced_version", (None, None))[0] if encoder_only else
                 (self._param_version() != getattr(self, "_synced_version", (None, None))[1] or getattr(self, "_params_stepped", False)))
        if self._dirty or self._handle is None or stale:
            self.sync_weights()
        if train_path:
            return self._forward_train(source, src_key_padding_mask, **kwargs)
        dev = self.device
        B, C, D, H, W = source.shape
        assert C == 1, "More than one channel"             # dino.py:14 / the (b d c) flatten at :125
        assert H % 14 == 0 and W % 14 == 0, \
            f"Input image height {H} / width {W} is not a multiple of patch size 14"   # patch_embed.py:72-73
        L = _cabi.lib()
        E, heads = self.encoder.embed_dim, self.encoder.num_heads
        N = (H // 14) * (W // 14) + 1 + self.num_registers
        transformer = self.slice_fusion_type == 'transformer'
        if save_attn and not transformer:
            # the reference's register_hooks walks self.slice_fusion.named_modules() (dino.py:257), which does not exist
            raise AttributeError(f"'{type(self).__name__}' object has no attribute 'slice_fusion'")
        full_maps = kwargs.get("_full_maps", None)
        tta = bool(kwargs.get("_tta", False))      # run_pred(use_tta=True): the 8 flipped variants as one batch (C ABI: tta)
        V = 8 if tta else 1
        feat_dim = self.emb_ch * D if self.slice_fusion_type == 'linear' else self.emb_ch
        mask = None
        if src_key_padding_mask is not None:               # dino.py:147-150
            mask = src_key_padding_mask.to(dev).to(torch.uint8).contiguous()
            if tuple(mask.shape) != (B, D):
                raise ValueError(f"src_key_padding_mask must be [B, D] = {(B, D)}, got {tuple(mask.shape)}")
        # Volume element type on the wire: the bf16 path rounds every voxel to bf16 before the patch GEMM, so a bf16 (or fp16)
        # `source` is taken as it is -- half the host-to-device bytes, bit-identical results for bf16; fp32 mode takes fp32.
        if self.precision == 'bf16' and source.dtype in (torch.bfloat16, torch.float16):
            src_dt = source.dtype
        else:
            src_dt = torch.float32
        src_code = _cabi.SRC_DTYPE[str(src_dt)]
        # `source.to(self.device)` (dino.py:121).  A host batch is moved in chunks of whole volumes on a copy stream so
        # that the H2D transfer of chunk k+1 overlaps the kernels of chunk k (volumes are independent: results are
        # bit-identical to a single call, tests/test_gpu_parity.py::test_batch_composition...).
        sched = self.h2d_chunk_volumes
        if isinstance(sched, int):
            sched = (sched,)
        chunks = [B]
        if source.device.type == "cpu" and sched and B > sched[0] and not tta:
            chunks, left, i = [], B, 0
            while left > 0:
                c = min(left, sched[min(i, len(sched) - 1)])
                chunks.append(c)
                left -= c
                i += 1
        chunk = max(chunks)
        want_enc = bool(kwargs.get("return_enc_cls", False))
        small = bool(self.graph_max_slices) and V * B * D <= self.graph_max_slices and full_maps is None
        if small:
            chunks, chunk = [B], B
        with torch.cuda.device(dev):
            need = _cabi.ctypes.c_size_t()
            if full_maps is not None:
                chunks, chunk = [B], B   # full maps are written for the whole batch in one call
            _cabi.check(L.mst_workspace_bytes(self._handle, V * chunk, D, H, W, _cabi.ctypes.byref(need)))
            if self._workspace is None or self._workspace.numel() < need.value or self._workspace.device != dev:
                self._workspace = None
                self._static = {}
                self._workspace = torch.empty(need.value, device=dev, dtype=torch.uint8)
                _cabi.check(L.mst_set_graph_threshold(self._handle, int(self.graph_max_slices or 0) * 400))

            def outputs():
                return (torch.empty((V * B, self.out_ch), device=dev, dtype=torch.float32) if self.enable_linear else None,
                        torch.empty((V * B, feat_dim), device=dev, dtype=torch.float32),
                        torch.empty((V * B * D, heads, N), device=dev, dtype=torch.float32) if save_attn else None,
                        torch.empty((V * B, synth.SLICE_HEADS, D + 1), device=dev, dtype=torch.float32) if save_attn else None,
                        torch.empty((V * B * D, E), device=dev, dtype=torch.float32) if want_enc else None)
            io = None
            if small:   # persistent buffers: the same argument tuple every call, so the library replays a captured graph
                skey = (B, D, H, W, bool(save_attn), src_dt, tta, want_enc, mask is not None)
                io = self._static.get(skey)
                if io is None:
                    if len(self._static) >= 8:
                        self._static.pop(next(iter(self._static)))
                    io = {"x": torch.empty((B, 1, D, H, W), device=dev, dtype=src_dt), "out": outputs(),
                          "mask": torch.empty((B, D), device=dev, dtype=torch.uint8) if mask is not None else None}
                    self._static[skey] = io
                logits, feat, plane, slc, enc = io["out"]
                if mask is not None:
                    io["mask"].copy_(mask)
                    mask = io["mask"]
            else:
                logits, feat, plane, slc, enc = outputs()
            cur = torch.cuda.current_stream()
            stream = _cabi.ctypes.c_void_p(cur.cuda_stream)

            def run(xc, b0, nb):
                sl = lambda t, per: None if t is None else t[b0 * per:(b0 + nb) * per]
                if tta:
                    sl = lambda t, per: t          # one call covers the batch; outputs are variant-major over all of it
                _cabi.check(L.mst_forward(self._handle, _cabi.ptr(xc), src_code, nb, D, H, W, _cabi.ptr(sl(mask, 1)), int(tta),
                                          _cabi.ptr(sl(logits, 1)), _cabi.ptr(sl(feat, 1)), _cabi.ptr(sl(enc, D)),
                                          _cabi.ptr(sl(plane, D)), _cabi.ptr(sl(slc, 1)), _cabi.ptr(full_maps),
                                          _cabi.ptr(self._workspace), self._workspace.numel(), stream))

            if io is not None:
                io["x"].copy_(source.reshape(B, 1, D, H, W), non_blocking=True)   # H2D (or D2D) + dtype conversion in one copy
                run(io["x"], 0, B)
                # the persistent buffers are overwritten by the next forward of this shape: hand out copies
                logits, feat, plane, slc, enc = (None if t is None else t.clone() for t in (logits, feat, plane, slc, enc))
            elif len(chunks) == 1:
                x = source.to(dev).to(src_dt).contiguous()
                run(x, 0, B)
            else:
                src = source if source.dtype == src_dt else source.to(src_dt)
                if (self._h2d is None or self._h2d[0].shape[1:] != (1, D, H, W) or self._h2d[0].device != dev
                        or self._h2d[0].shape[0] < chunk or self._h2d[0].dtype != src_dt):
                    self._h2d = None
                    self._h2d = [torch.empty((chunk, 1, D, H, W), device=dev, dtype=src_dt) for _ in range(2)]
                    self._copy_stream = torch.cuda.Stream(device=dev)
                copied = [torch.cuda.Event() for _ in range(2)]
                freed = [torch.cuda.Event() for _ in range(2)]
                self._copy_stream.wait_stream(cur)   # buffers may still be read by a previous forward
                b0 = 0
                for k, nb in enumerate(chunks):
                    buf = self._h2d[k & 1]
                    with torch.cuda.stream(self._copy_stream):
                        if k >= 2:
                            self._copy_stream.wait_event(freed[k & 1])
                        buf[:nb].copy_(src[b0:b0 + nb], non_blocking=True)
                        copied[k & 1].record(self._copy_stream)
                    cur.wait_event(copied[k & 1])
                    run(buf, b0, nb)
                    freed[k & 1].record(cur)
                    b0 += nb
        if save_attn:
            # The reference keeps 12 x [BD,heads,N,N] (dino.py:241); its getters only ever read row 0 of the last
            # one.  We keep that row, shaped [BD,heads,1,N] so that `attention_maps[-1][:, :, 0, 1:]` still works.
            self.attention_maps = [plane.unsqueeze(2)]
            self.attention_maps_slice = [slc.unsqueeze(2)]
            self._last = (B, D, H, W)
            self._last_tta = tta
            # get_attention_cls needs every block's full map: recomputed on demand from these (caller-owned) inputs
            self._last_inputs = (source, src_key_padding_mask)
        self._enc_cls = enc
        if kwargs.get('without_linear', False) or not self.enable_linear:   # dino.py:164-165; nn.Identity head (:103)
            return feat
        return logits

    # -- training step on a frozen encoder (BASELINE config 5, frozen-encoder slice; new_vit_b200/training.py) -----------
    def _forward_train(self, source, src_key_padding_mask=None, **kwargs):
        """train()-mode forward with autograd: the encoder runs as in inference (it must be frozen: its backward pass is not
        built), the slice transformer + head run in the differentiable CUDA path.  Dropouts are 0 and drop_path is 0 in the
        reference (dino.py:89), so train-mode arithmetic equals eval-mode arithmetic."""
        train_encoder = any(p.requires_grad for p in self.encoder.parameters())
        if train_encoder and (self.precision != 'bf16' or self.num_registers or any(k.endswith("ls1.gamma") for k in self.state_dict())):
            raise NotImplementedError(
                "the encoder's backward pass is built for precision='bf16' on the vendored-factory architecture (no LayerScale, no "
                "registers): construct with freeze=True (dino.py:69-71) to train the slice transformer + head only")
        if (self.slice_fusion_type != 'transformer' or hasattr(self, "bottleneck") or hasattr(self, "slice_pos_emb")
                or self.rotary is not None or not self.enable_linear):
            raise NotImplementedError("the differentiable slice path covers the default construction (dino.py:84-103: transformer "
                                      "fusion, no bottleneck / slice position embedding / rotary, linear head)")
        B, C, D, H, W = source.shape
        if train_encoder:
            # every parameter trains (main_train.py:110-126 on the default construction): the encoder's CUDA training forward keeps
            # its activations for the CUDA backward pass; all tensors of encoder.* take part, in state_dict order
            named = [(n, p) for n, p in self.named_parameters() if n.startswith("encoder.")]
            enc = EncoderFunction.apply(self, source, tuple(n for n, _ in named), *[p for _, p in named]).view(B, D, -1)
        else:
            was_training = self.training
            self.training = False                   # (re-enters forward() on the inference path; submodules are parameter holders)
            try:
                with torch.no_grad():
                    self.forward(source, src_key_padding_mask=src_key_padding_mask, return_enc_cls=True, _encoder_only=True)
            finally:
                self.training = was_training
            enc = self._enc_cls.view(B, D, -1)
        mask = None
        if src_key_padding_mask is not None:
            mask = src_key_padding_mask.to(self.device).to(torch.uint8).contiguous()
        sd = dict(self.named_parameters())
        params = [sd[n] for n in SLICE_PARAM_NAMES]
        return SliceHeadFunction.apply(self._handle, enc, mask, synth.SLICE_HEADS, train_encoder, *params)

    # -- instrumentation ------------------------------------------------------------------------------------
    def launch_count(self):
        """CUDA kernels launched by this model's handle so far (kernels inside a replayed CUDA graph are counted)."""
        return int(_cabi.lib().mst_launch_count(self._handle)) if self._handle is not None else 0

    def graph_replays(self):
        """Forwards that ran as a replayed CUDA graph (small batches, see graph_max_slices)."""
        return int(_cabi.lib().mst_graph_replays(self._handle)) if self._handle is not None else 0

    def profile_begin(self):
        if self._dirty or self._handle is None:
            self.sync_weights()
        _cabi.check(_cabi.lib().mst_profile_begin(self._handle))

    def profile_end(self):
        """{category: (milliseconds, launches)} accumulated since profile_begin (synchronises the device)."""
        ms = (_cabi.ctypes.c_double * 32)()
        n = (_cabi.ctypes.c_int64 * 32)()
        _cabi.check(_cabi.lib().mst_profile_end(self._handle, ms, n, 32))
        names = _cabi.lib().mst_profile_categories().decode().split(",")
        return {k: (ms[i], int(n[i])) for i, k in enumerate(names)}

    # -- getters (dino.py:173-212) ---------------------------------------------------------------------
    def _saliency(self, want_maps=False, want_plane=False, want_slice=False, want_coarse=False, want_full=False, size=None):
        if not self.attention_maps or not self.attention_maps_slice:
            raise IndexError("list index out of range")    # what the reference raises before a save_attn forward
        plane = self.attention_maps[-1][:, :, 0, :].contiguous()
        slc = self.attention_maps_slice[-1][:, :, 0, :].contiguous()
        dev = plane.device
        BD, heads, N = plane.shape
        B, sheads, L = slc.shape
        D = L - 1
        tta = bool(getattr(self, "_last_tta", False))
        if tta:                                            # the stored rows hold 8 flipped variants per volume
            if want_maps or want_plane:
                raise MSTError("get_attention_maps / get_plane_attention are per variant; after a TTA forward use saliency_volume()")
            B, BD = B // 8, BD // 8
        skip = 1 + self.num_registers                      # dino.py:191
        P = N - skip
        if self._last is not None and self._last[0] * self._last[1] == BD:
            H, W = self._last[2], self._last[3]
        else:
            g = int(P ** 0.5)
            H = W = 14 * g
        gh, gw = H // 14, W // 14
        if size is not None:
            H, W = size
        with torch.cuda.device(dev):
            maps = torch.empty((BD, heads, P), device=dev, dtype=torch.float32) if want_maps else None
            pl = torch.empty((BD, heads, P), device=dev, dtype=torch.float32) if want_plane else None
            sl = torch.empty((BD,), device=dev, dtype=torch.float32) if want_slice else None
            coarse = torch.empty((B, 1, D, gh, gw), device=dev, dtype=torch.float32) if (want_coarse or want_full) else None
            full = torch.empty((B, 1, D, H, W), device=dev, dtype=torch.float32) if want_full else None
            stream = _cabi.ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _cabi.check(_cabi.lib().mst_saliency(self._handle, _cabi.ptr(plane), _cabi.ptr(slc), B, D, heads, sheads, skip, gh, gw,
                                                 H, W, int(tta), _cabi.ptr(maps), _cabi.ptr(pl), _cabi.ptr(sl), _cabi.ptr(coarse),
                                                 _cabi.ptr(full), stream))
        return maps, pl, sl, coarse, full

    def get_slice_attention(self):
        """[B*D, 1, 1] (dino.py:173-187)."""
        return self._saliency(want_slice=True)[2][:, None, None]

    def get_plane_attention(self):
        """[B*D, heads, P]: CLS->patch attention of the last block, patch 0 zeroed, renormalised (dino.py:189-195)."""
        return self._saliency(want_plane=True)[1]

    def get_attention_maps(self):
        """[B*D, heads, P] = slice attention x plane attention (dino.py:197-202)."""
        return self._saliency(want_maps=True)[0]

    def get_attention_cls(self):
        """Attention rollout over all encoder blocks (dino.py:204-212) -> [B*D, heads, N, N].

        The reference keeps every block's [B*D,heads,N,N] map from the save_attn forward (dino.py:241; 39 GB at
        256 volumes); here only row 0 of the last one is kept, and the full maps are recomputed on demand from the
        inputs of the last save_attn forward, then multiplied right to left on the device."""
        if not self.attention_maps or getattr(self, "_last_inputs", None) is None:
            raise IndexError("list index out of range")
        source, mask = self._last_inputs
        B, D, H, W = self._last
        heads, depth = self.encoder.num_heads, self.encoder.depth
        N = (H // 14) * (W // 14) + 1 + self.num_registers
        dev = self.device
        need = (depth + 2) * B * D * heads * N * N * 4
        free = torch.cuda.mem_get_info(dev)[0]
        if need > free:
            raise MSTError(f"get_attention_cls needs {need / 2**30:.1f} GiB for {depth} full maps of {B * D} slices "
                           f"({free / 2**30:.1f} GiB free): call it on a smaller batch")
        with torch.cuda.device(dev):
            maps = torch.empty((depth, B * D, heads, N, N), device=dev, dtype=torch.float32)
            with torch.no_grad():
                self.forward(source, save_attn=False, src_key_padding_mask=mask, _full_maps=maps)
            out = torch.empty((B * D, heads, N, N), device=dev, dtype=torch.float32)
            scratch = torch.empty_like(out)
            stream = _cabi.ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _cabi.check(_cabi.lib().mst_rollout(_cabi.ptr(maps), depth, B * D * heads, N, _cabi.ptr(out), _cabi.ptr(scratch), stream))
        return out

    def interpolated_pos_embed(self, H, W):
        """`encoder.interpolate_pos_encoding` (vision_transformer.py:179-211) for an H x W input -> [1, 1+P, E]."""
        if self._dirty or self._handle is None:
            self.sync_weights()
        dev = self.device
        with torch.cuda.device(dev):
            out = torch.empty((1, 1 + (H // 14) * (W // 14), self.encoder.embed_dim), device=dev, dtype=torch.float32)
            stream = _cabi.ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _cabi.check(_cabi.lib().mst_pos_embed(self._handle, H, W, _cabi.ptr(out), stream))
        return out

    def saliency_volume(self, size=None):
        """Batched form of main_predict.py:73-74,93-105,161-162: returns
        (weight [B,1,D,H,W] upsampled map, weight_slice [B,1,D,1,1] slice weights)."""
        _, _, sl, coarse, full = self._saliency(want_slice=True, want_full=True, size=size)
        B, _, D = coarse.shape[:3]
        return full, sl.view(B, 1, D, 1, 1)


def quantile(x, q):
    """np.quantile(x[i], q) for every item i of a CUDA fp32 tensor [items, ...] (numpy 'linear' method), on the device:
    the 0.995/0.999 clip and 0.999 threshold of scripts/main_predict.py:243-245,296.  Returns float64 [items, len(q)]."""
    if x.device.type != "cuda":
        raise MSTError("quantile runs on a CUDA tensor only (no CPU fallback)")
    qs = [float(v) for v in (q if isinstance(q, (list, tuple)) else [q])]
    if not all(0.0 <= v <= 1.0 for v in qs):
        raise ValueError("Quantiles must be in the range [0, 1]")   # numpy's message
    x = x.detach().to(torch.float32).contiguous()
    items = x.shape[0]
    n = x.numel() // items
    L = _cabi.lib()
    with torch.cuda.device(x.device):
        qd = torch.tensor(qs, dtype=torch.float64, device=x.device)
        out = torch.empty((items, len(qs)), dtype=torch.float64, device=x.device)
        need = _cabi.ctypes.c_size_t()
        _cabi.check(L.mst_quantile_workspace_bytes(items, len(qs), _cabi.ctypes.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device=x.device)
        stream = _cabi.ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _cabi.check(L.mst_quantile(_cabi.ptr(x), n, items, _cabi.ptr(qd), len(qs), _cabi.ptr(out), _cabi.ptr(ws), need.value, stream))
    return out


def _pred_trans(model, source, src_key_padding_mask, save_attn=False, use_softmax=True, _tta=False):
    """scripts/main_predict.py:55-106, generalised from batch 1 to batch B.  With _tta the forward covers the 8 flipped
    variants as one batch (predictions come back [8, B, out_ch]) and the saliency kernels un-flip and average them."""
    with torch.no_grad():
        pred = model(source, src_key_padding_mask=src_key_padding_mask, save_attn=save_attn, _tta=_tta)
    if use_softmax:
        pred = torch.softmax(pred, dim=-1)
    if _tta:
        pred = pred.view(8, source.shape[0], -1)
    if not save_attn:
        return pred, None, None
    weight, weight_slice = model.saliency_volume(size=tuple(source.shape[3:]))
    weight_slice = weight_slice.expand(*source.shape)
    return pred, weight, weight_slice


def run_pred(model, batch, save_attn=False, use_softmax=True, use_tta=False):
    """scripts/main_predict.py:133-164.  use_tta: the script's eight forwards (:147-158) run as ONE forward over the 8 flipped
    variants (flips are index arithmetic on the patch load), the coarse maps are un-flipped and averaged in the script's
    summation order inside the combine kernel, and the x14 upsample (F.interpolate trilinear, :161-162) runs once on the
    average: one forward + two map launches.  The un-flipped padding mask goes to every variant, as the script passes it (:149)."""
    source, mask = batch['source'], batch.get('src_key_padding_mask', None)
    pred, weight, weight_slice = _pred_trans(model, source, mask, save_attn, use_softmax, _tta=use_tta)
    if use_tta:
        p = pred[0]
        for i in range(1, 8):      # the script's own order of additions (:151), on [B, out_ch] values
            p = p + pred[i]
        pred = p / 8
    return pred, weight, weight_slice
