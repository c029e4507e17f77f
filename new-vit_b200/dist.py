"""Volume sharding across the GPUs of one box.

Volumes are independent units (slices interact only inside their own volume's slice transformer, reference
dino.py:138-153), so rank r processes a contiguous block of whole volumes with replicated weights; the forward has
no data-path collective.  Only the `[B_local, out_ch]` logits (and, on request, the coarse `[B_local, D, g, g]`
saliency maps) are gathered; full-resolution maps stay sharded.  Backend: NCCL on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_volumes(total, rank, world):
    """Contiguous block [start, start+count) of volumes for `rank`; the first `total % world` ranks get one more."""
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def gather_volumes(local, total, group=None):
    """All-gather per-volume results `local` [count_r, ...] into [total, ...] in global volume order on every rank.
    Uneven shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert local.shape[0] == total
        return local
    world = dist.get_world_size(group)
    counts = [shard_volumes(total, r, world)[1] for r in range(world)]
    cmax = max(counts)
    pad = local.new_zeros((cmax,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def predict_sharded(model, source_all, src_key_padding_mask=None, save_attn=False, group=None):
    """Every rank holds (or can index) the full host batch; it runs its own block of volumes through `model` and the
    logits are gathered.  Returns (logits [total, out_ch] on every rank, (start, count) of this rank's shard)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    total = source_all.shape[0]
    start, count = shard_volumes(total, rank, world)
    mask = None if src_key_padding_mask is None else src_key_padding_mask[start:start + count]
    if count == 0:   # more ranks than volumes: this rank contributes an empty block (the forward rejects an empty batch)
        out_ch = getattr(model, "out_ch", None) or getattr(model, "emb_ch")
        local = torch.empty((0, out_ch), dtype=torch.float32, device=getattr(model, "device", source_all.device))
    else:
        with torch.no_grad():
            local = model(source_all[start:start + count], save_attn=save_attn, src_key_padding_mask=mask)
    return gather_volumes(local, total, group), (start, count)
