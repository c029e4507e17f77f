"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: volume sharding and logit gathering."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from new_vit_b200 import dist as mdist


def test_shard_volumes_partitions_exactly():
    for total in (1, 2, 7, 64, 255, 256):
        for world in (1, 2, 3, 4, 8):
            blocks = [mdist.shard_volumes(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


class _FakeModel(torch.nn.Module):
    """Stands in for the CUDA model on CPU: one 'logit' pair per volume computed from the volume itself."""
    out_ch = 2     # read by predict_sharded on a rank whose shard is empty

    def forward(self, source, save_attn=False, src_key_padding_mask=None):
        s = source.reshape(source.shape[0], -1)
        out = torch.stack([s.sum(1), s.mean(1)], dim=1)
        if src_key_padding_mask is not None:
            out = out + src_key_padding_mask.float().sum(1, keepdim=True)
        return out


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(total, 1, 2, 4, 4, generator=g)
        mask = torch.zeros(total, 2, dtype=torch.bool)
        mask[1::2, 1] = True
        logits, (start, count) = mdist.predict_sharded(_FakeModel(), x, mask)
        ref = _FakeModel()(x, src_key_padding_mask=mask)
        ok = torch.allclose(logits, ref) and logits.shape == (total, 2)
        q.put((rank, bool(ok), start, count))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 5, 1])   # 1: the second rank's shard is empty
def test_predict_sharded_gloo_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert res[0][2] == 0 and res[0][2] + res[0][3] == res[1][2] and res[1][2] + res[1][3] == total
