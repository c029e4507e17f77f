"""Training-side surface of the drop-in (reference mst/models/base_model.py): the Lightning-module methods the training script
drives (`training_step` / `validation_step` / `_step` / `compute_loss` / `configure_optimizers`, base_model.py:24-48,103-110,
148-181) and the training step of BASELINE.json config 5 for the frozen-encoder construction (`freeze=True`, dino.py:69-71):

    encoder (frozen)  ->  CUDA forward as in inference, per-slice features [B, D, E]
    slice transformer + head  ->  mst_slice_train_forward / mst_slice_train_backward (csrc/train.cu) behind a torch.autograd.Function
    loss  ->  the caller's (CrossEntropyLoss on [B, out_ch], base_model.py:159,180-181)
    optimizer  ->  FusedAdamW: torch.optim.AdamW's update as one kernel over a flat buffer (mst_adamw); with torch.distributed
                   initialised the flat gradient is all-reduced (NCCL) first, which is what Lightning's DDP does for the reference

pytorch_lightning and torchmetrics are not installed here, so the Lightning plumbing (`self.log`, metric objects) is restated with
the same names and call conventions; nothing below runs on the CPU except bookkeeping."""
import ctypes

import torch
import torch.nn as nn

from . import _cabi
from ._cabi import MSTError

# order of the 17 tensors in the C ABI (include/mst_b200.h, mst_slice_train_forward)
SLICE_PARAM_NAMES = (
    "cls_token",
    "slice_fusion.layers.0.norm1.weight", "slice_fusion.layers.0.norm1.bias",
    "slice_fusion.layers.0.self_attn.in_proj_weight", "slice_fusion.layers.0.self_attn.in_proj_bias",
    "slice_fusion.layers.0.self_attn.out_proj.weight", "slice_fusion.layers.0.self_attn.out_proj.bias",
    "slice_fusion.layers.0.norm2.weight", "slice_fusion.layers.0.norm2.bias",
    "slice_fusion.layers.0.linear1.weight", "slice_fusion.layers.0.linear1.bias",
    "slice_fusion.layers.0.linear2.weight", "slice_fusion.layers.0.linear2.bias",
    "slice_fusion.norm.weight", "slice_fusion.norm.bias",
    "linear.weight", "linear.bias",
)


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class SliceHeadFunction(torch.autograd.Function):
    """logits = head(slice_transformer([cls_token; enc])[:, 0]) with the CUDA forward / backward of csrc/train.cu."""

    @staticmethod
    def forward(ctx, handle, enc, mask, heads, want_denc, *params):
        B, D, E = enc.shape
        C = params[-1].shape[0]
        L = _cabi.lib()
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != enc.device:
                raise MSTError("slice training: parameters must be contiguous fp32 tensors on the model's device")
        sb, fb = ctypes.c_size_t(), ctypes.c_size_t()
        _cabi.check(L.mst_slice_train_bytes(B, D, E, heads, C, ctypes.byref(sb), ctypes.byref(fb)))
        with torch.cuda.device(enc.device):
            saved = torch.empty(sb.value // 4, device=enc.device, dtype=torch.float32)
            logits = torch.empty((B, C), device=enc.device, dtype=torch.float32)
            _cabi.check(L.mst_slice_train_forward(handle, _cabi.ptr(enc), _cabi.ptr(mask), _ptr_array(params), B, D, E, heads, C,
                                                  _cabi.ptr(saved), _cabi.ptr(logits), _stream()))
        ctx.handle, ctx.heads, ctx.fbytes, ctx.want_denc = handle, heads, fb.value, want_denc
        ctx.save_for_backward(enc, saved, *params)
        ctx.mask = mask
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        enc, saved, *params = ctx.saved_tensors
        B, D, E = enc.shape
        C = params[-1].shape[0]
        L = _cabi.lib()
        with torch.cuda.device(enc.device):
            dl = dlogits.contiguous().float()
            factors = torch.empty(ctx.fbytes // 4, device=enc.device, dtype=torch.float32)
            grads = [torch.empty_like(p) for p in params]
            denc = torch.empty_like(enc) if ctx.want_denc else None
            _cabi.check(L.mst_slice_train_backward(ctx.handle, _cabi.ptr(enc), _cabi.ptr(dl), _ptr_array(params), _cabi.ptr(saved),
                                                   _cabi.ptr(factors), _ptr_array(grads), _cabi.ptr(denc), B, D, E, ctx.heads, C, _stream()))
        return (None, denc, None, None, None, *grads)


class EncoderFunction(torch.autograd.Function):
    """enc_cls [B*D, E] = DINOv2 encoder(source) with the CUDA training forward (every block's activations kept in `workspace`) and
    the CUDA backward pass (csrc/train_enc.cu, api.cu train_forward / train_backward).  `params` are the encoder's parameters in
    `names` order; their gradients come back as fp32 tensors written by mst_train_backward."""

    @staticmethod
    def forward(ctx, model, source, names, *params):
        L = _cabi.lib()
        dev = model.device
        B, C, D, H, W = source.shape
        E = model.encoder.embed_dim
        if model.precision == 'bf16' and source.dtype in (torch.bfloat16, torch.float16):
            src_dt = source.dtype
        else:
            src_dt = torch.float32
        with torch.cuda.device(dev):
            x = source.to(dev).to(src_dt).contiguous()
            need = ctypes.c_size_t()
            _cabi.check(L.mst_train_workspace_bytes(model._handle, B, D, H, W, ctypes.byref(need)))
            ws = getattr(model, "_train_workspace", None)
            if ws is None or ws.numel() < need.value or ws.device != dev:
                model._train_workspace = None
                ws = model._train_workspace = torch.empty(need.value, device=dev, dtype=torch.uint8)
            enc = torch.empty((B * D, E), device=dev, dtype=torch.float32)
            _cabi.check(L.mst_train_forward(model._handle, _cabi.ptr(x), _cabi.SRC_DTYPE[str(src_dt)], B, D, H, W, _cabi.ptr(enc), _cabi.ptr(ws),
                                            ws.numel(), _stream()))
        ctx.model, ctx.shape, ctx.names, ctx.ws = model, (B, D, H, W), names, ws
        ctx.save_for_backward(*params)
        return enc

    @staticmethod
    def backward(ctx, denc):
        L = _cabi.lib()
        model, (B, D, H, W), params = ctx.model, ctx.shape, ctx.saved_tensors
        with torch.cuda.device(model.device):
            # One flat fp32 buffer receives every gradient of the encoder; it and its registration with the library (mst_set_grad)
            # are kept across steps (167 allocations and ctypes calls per step otherwise).
            cache = getattr(model, "_enc_grad_cache", None)
            key = (tuple(ctx.names), tuple(tuple(p.shape) for p in params), str(model.device), getattr(model._handle, "value", id(model._handle)))
            if cache is None or cache["key"] != key:
                sizes = [(p.numel() + 3) // 4 * 4 for p in params]
                flat = torch.empty(sum(sizes), device=model.device, dtype=torch.float32)
                views, off = [], 0
                for n, p, sz in zip(ctx.names, params, sizes):
                    if n == "encoder.mask_token":          # never read by the path (vision_transformer.py:216 masks=None)
                        views.append(None)
                    else:
                        g = flat[off:off + p.numel()].view(p.shape)
                        _cabi.check(L.mst_set_grad(model._handle, n.encode(), _cabi.ptr(g), g.numel()))
                        views.append(g)
                    off += sz
                cache = model._enc_grad_cache = {"key": key, "flat": flat, "views": views}
            views = cache["views"]
            _cabi.check(L.mst_train_backward(model._handle, _cabi.ptr(denc.contiguous().float()), B, D, H, W, _cabi.ptr(ctx.ws), ctx.ws.numel(),
                                             _stream()))
            # With .grad already allocated (FusedAdamW keeps them as views of its flat buffer) the gradients are accumulated by ONE
            # multi-tensor add instead of 167 AccumulateGrad launches; otherwise autograd gets them as usual (copies: the flat buffer
            # is overwritten by the next backward pass).
            named = dict(model.named_parameters())
            live = [(named[n], g) for n, g in zip(ctx.names, views) if g is not None]
            if all(p.grad is not None and p.grad.dtype == torch.float32 and p.grad.device == g.device for p, g in live):
                torch._foreach_add_([p.grad for p, _ in live], [g for _, g in live])
                return (None, None, None) + (None,) * len(views)
        return (None, None, None, *[None if g is None else g.clone() for g in views])


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW's update (decoupled weight decay, bias correction, eps outside the sqrt) as ONE kernel over a flat
    fp32 buffer per parameter group (mst_adamw).  The parameters are re-pointed at views of the flat buffer, so `.grad` written by
    autograd lands in a flat gradient buffer too; with torch.distributed initialised that buffer is all-reduced before the
    update (sum over ranks, scaled by 1/world inside the kernel) -- the gradient synchronisation Lightning's DDP performs for the
    reference (main_train.py:110-123 with devices='auto')."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, handle=None, all_reduce=True, on_step=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._handle, self._all_reduce, self._on_step = handle, all_reduce, on_step
        self._flat = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad]
            if not ps:
                self._flat.append(None)
                continue
            dev = ps[0].device
            if dev.type != "cuda" or any(p.dtype != torch.float32 or p.device != dev for p in ps):
                raise MSTError("FusedAdamW updates fp32 CUDA parameters of one device (no CPU fallback)")
            sizes = [(p.numel() + 3) // 4 * 4 for p in ps]      # every tensor starts on a 16-byte boundary
            n = sum(sizes)
            flat_p = torch.zeros(n, device=dev)
            flat_g = torch.zeros(n, device=dev)
            off, offsets = 0, []
            for p, sz in zip(ps, sizes):
                offsets.append(off)
                flat_p[off:off + p.numel()].copy_(p.data.reshape(-1))
                p.data = flat_p[off:off + p.numel()].view_as(p)
                p.grad = flat_g[off:off + p.numel()].view_as(p)
                off += sz
            self._flat.append(dict(p=flat_p, g=flat_g, m=torch.zeros(n, device=dev), v=torch.zeros(n, device=dev), step=0, params=ps, offsets=offsets))

    def zero_grad(self, set_to_none=False):
        for f in self._flat:
            if f is not None:
                f["g"].zero_()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        L = _cabi.lib()
        import torch.distributed as dist
        world = dist.get_world_size() if (self._all_reduce and dist.is_available() and dist.is_initialized()) else 1
        for group, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            for p, off in zip(f["params"], f["offsets"]):   # a .grad that autograd (re)created outside the flat buffer is folded back
                want = f["g"].data_ptr() + 4 * off
                if p.grad is not None and p.grad.data_ptr() != want:
                    f["g"][off:off + p.numel()].copy_(p.grad.reshape(-1))
                    p.grad = f["g"][off:off + p.numel()].view_as(p)
            if world > 1:
                dist.all_reduce(f["g"])                        # NCCL sum over the ranks; 1/world is applied inside the kernel
            f["step"] += 1
            b1, b2 = group["betas"]
            with torch.cuda.device(f["p"].device):
                _cabi.check(L.mst_adamw(self._handle, _cabi.ptr(f["p"]), _cabi.ptr(f["g"]), _cabi.ptr(f["m"]), _cabi.ptr(f["v"]),
                                        f["p"].numel(), group["lr"], b1, b2, group["eps"], group["weight_decay"], f["step"],
                                        1.0 / world, _stream()))
        if self._on_step is not None:
            self._on_step()     # the kernel wrote the parameters behind torch's back (no version bump): tell the owner
        return loss


class _Accuracy:
    """torchmetrics.Accuracy(task='multiclass', num_classes) as base_model.py:145,164,176 uses it: micro accuracy over an epoch."""

    def __init__(self, num_classes):
        self.num_classes = num_classes
        self.reset()

    def reset(self):
        self.correct, self.total = 0, 0

    def update(self, pred, target):
        self.correct += int((pred.argmax(-1) == target).sum())
        self.total += int(target.numel())

    def compute(self):
        return torch.tensor(self.correct / max(self.total, 1))


class _AUROC:
    """torchmetrics.AUROC(task='multiclass', num_classes): macro average of the one-vs-rest areas (base_model.py:144,165,176)."""

    def __init__(self, num_classes):
        self.num_classes = num_classes
        self.reset()

    def reset(self):
        self.preds, self.targets = [], []

    def update(self, pred, target):
        self.preds.append(torch.softmax(pred.detach().float(), -1).cpu())
        self.targets.append(target.detach().cpu())

    def compute(self):
        if not self.preds:
            return torch.tensor(0.0)
        p, t = torch.cat(self.preds), torch.cat(self.targets)
        areas = []
        for c in range(self.num_classes):
            pos, neg = p[t == c, c], p[t != c, c]
            if len(pos) == 0 or len(neg) == 0:
                continue
            gt = (pos[:, None] > neg[None, :]).float().mean() + 0.5 * (pos[:, None] == neg[None, :]).float().mean()
            areas.append(gt)
        return torch.stack(areas).mean() if areas else torch.tensor(0.0)


class LightningSurface:
    """Mixin with the methods of VeryBasicModel / BasicModel / BasicClassifier (base_model.py:10-181) that main_train.py and
    Lightning's loop call on the model.  State it needs (`loss_func`, `optimizer`, `optimizer_kwargs`, metric dictionaries,
    step counters) is created by `_init_lightning_surface`."""

    def _init_lightning_surface(self, out_ch, loss=nn.CrossEntropyLoss, loss_kwargs=None, optimizer=None, optimizer_kwargs=None,
                                lr_scheduler=None, lr_scheduler_kwargs=None):
        self._step_train = self._step_val = self._step_test = -1            # base_model.py:15-17
        self.loss_kwargs = dict(loss_kwargs or {})
        self.loss_func = loss(**self.loss_kwargs)                           # base_model.py:139
        self.optimizer = optimizer or torch.optim.AdamW                     # base_model.py:124
        self.optimizer_kwargs = optimizer_kwargs if optimizer_kwargs is not None else {'lr': 1e-4, 'weight_decay': 1e-2}
        self.lr_scheduler, self.lr_scheduler_kwargs = lr_scheduler, dict(lr_scheduler_kwargs or {})
        self.auc_roc = {s: _AUROC(out_ch) for s in ("train_", "val_", "test_")}   # base_model.py:144-145 ('train' is not a legal key)
        self.acc = {s: _Accuracy(out_ch) for s in ("train_", "val_", "test_")}
        self.logged = {}
        self.batch_size = None

    # -- Lightning's logger hook: keeps the last value per key (there is no Trainer here) --
    def log(self, name, value, batch_size=None, on_step=None, on_epoch=None, sync_dist=False, **kwargs):
        v = value.detach() if isinstance(value, torch.Tensor) else torch.as_tensor(value)
        if sync_dist:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                v = v.to(self.device).float().clone()
                dist.all_reduce(v)
                v /= dist.get_world_size()
        self.logged[name] = v

    def training_step(self, batch, batch_idx):                              # base_model.py:29-31
        self._step_train += 1
        return self._step(batch, batch_idx, "train", self._step_train)

    def validation_step(self, batch, batch_idx):                            # base_model.py:33-35
        self._step_val += 1
        return self._step(batch, batch_idx, "val", self._step_val)

    def test_step(self, batch, batch_idx):                                  # base_model.py:37-39
        self._step_test += 1
        return self._step(batch, batch_idx, "test", self._step_test)

    def on_train_epoch_end(self):
        self._epoch_end("train")

    def on_validation_epoch_end(self):
        self._epoch_end("val")

    def on_test_epoch_end(self, outputs=None):                              # (stale `outputs` argument as in base_model.py:47)
        self._epoch_end("test")

    def _step(self, batch, batch_idx, state, step):                         # base_model.py:148-170
        target = batch['target']
        batch_size = target.shape[0]
        self.batch_size = batch_size
        pred = self(**batch)                                                # uid / target land in forward's **kwargs
        target = target.to(pred.device)
        logging_dict = {'loss': self.compute_loss(pred, target)}
        with torch.no_grad():
            self.acc[state + "_"].update(pred, target)
            self.auc_roc[state + "_"].update(pred, target)
            for metric_name, metric_val in logging_dict.items():
                self.log(f"{state}/{metric_name}", metric_val, batch_size=batch_size, on_step=True, on_epoch=True, sync_dist=False)
        return logging_dict['loss']

    def _epoch_end(self, state):                                            # base_model.py:172-178
        for name, value in [("ACC", self.acc[state + "_"]), ("AUC_ROC", self.auc_roc[state + "_"])]:
            self.log(f"{state}/{name}", value.compute(), batch_size=self.batch_size, on_step=False, on_epoch=True, sync_dist=True)
            value.reset()

    def compute_loss(self, pred, target):                                   # base_model.py:180-181
        return self.loss_func(pred, target)

    def configure_optimizers(self):                                         # base_model.py:103-110
        params = [p for p in self.parameters()]
        if self.optimizer in (torch.optim.AdamW, FusedAdamW) and self.device.type == "cuda":
            # torch.optim.AdamW's update as one fused kernel over the trainable parameters (same arithmetic)
            kw = {k: v for k, v in self.optimizer_kwargs.items() if k in ("lr", "betas", "eps", "weight_decay")}
            optimizer = FusedAdamW([p for p in params if p.requires_grad], handle=getattr(self, "_handle", None),
                                   on_step=lambda: setattr(self, "_params_stepped", True), **kw)
        else:
            optimizer = self.optimizer(params, **self.optimizer_kwargs)
        if self.lr_scheduler is not None:
            lr_scheduler = self.lr_scheduler(optimizer, **self.lr_scheduler_kwargs)
            return [optimizer], [{"scheduler": lr_scheduler, "interval": "step", "frequency": 1}]
        return [optimizer]
