"""Input pipeline on the GPU (csrc/prep.cu through the C ABI, `new_vit_b200.duke_transform`) against
oracle/duke_transform_oracle.py on the same seeded volumes.  Floating point: the cutoffs are order
statistics blended in fp32 (tolerance 1e-6 relative), mean / std are fp64 sums here and fp32 cascade sums in ATen (1e-5
relative), the normalised voxels 2e-5 absolute (values are O(1)); min, max and the masked count are exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _volume(shape, seed, kind="gamma"):
    rng = np.random.default_rng(seed)
    if kind == "gamma":      # MRI-like: non-negative, long tail
        return rng.gamma(2.0, 100.0, size=shape).astype(np.float32)
    if kind == "ct":         # signed, with a large flat background (ties at the minimum)
        v = rng.normal(0.0, 300.0, size=shape).astype(np.float32)
        v[rng.random(shape) < 0.3] = -1024.0
        return v
    if kind == "geom":       # all values distinct and a factor 2^(1/8) apart: adjacent order statistics never share a 12-bit key
        n = int(np.prod(shape))   # prefix, so the below / above ranks of a percentile walk different radix buckets
        return (2.0 ** (rng.permutation(n) / 8.0 - 60.0)).astype(np.float32).reshape(shape)
    return np.round(rng.normal(0, 3, size=shape)).astype(np.float32)    # heavy ties everywhere


CASES = [
    # (W0, H0, D0), image_crop (W, H, D), kind
    ((256, 240, 40), (224, 224, 32), "gamma"),     # crop on every axis (odd/even amounts)
    ((224, 224, 32), (224, 224, 32), "ct"),        # nothing to crop or pad
    ((200, 210, 20), (224, 224, 32), "gamma"),     # pad on every axis (corners take the global minimum)
    ((230, 190, 32), (224, 224, 32), "ct"),        # crop W, pad H
    ((250, 224, 27), (224, 224, 32), "ties"),      # crop W, pad D by an odd amount
    ((61, 45, 9), (56, 56, 8), "gamma"),           # small ragged tiles (W, D not multiples of 32)
    ((33, 70, 50), (40, 64, 36), "ct"),            # pad W, crop H and D, D > 32 (two d-tiles)
    ((72, 60, 48), (56, 56, 40), "gamma"),         # 16-byte path on the full tiles, scalar path on the ragged ones (W, D)
    ((8, 8, 12), (8, 8, 12), "geom"),              # every selection under its own radix prefix
]


@pytest.mark.parametrize("shape,crop,kind", CASES)
def test_duke_transform_matches_oracle(shape, crop, kind):
    from new_vit_b200 import duke_transform
    from oracle import duke_transform_oracle as O
    v = _volume(shape, seed=sum(shape), kind=kind)
    want, st = O.duke_transform(v, image_crop=crop)
    got, stats = duke_transform(torch.from_numpy(v).cuda(), image_crop=crop, return_stats=True)
    assert got.shape == (1, 1, crop[2], crop[1], crop[0]) and got.dtype == torch.float32
    s = stats[0].cpu().numpy()
    assert s[0] == st["min"] and s[1] == st["max"] and int(s[6]) == st["count"] and s[7] == 0
    np.testing.assert_allclose(s[2:4], [st["lo"], st["hi"]], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(s[4:6], [st["mean"], st["std"]], rtol=1e-5)
    torch.testing.assert_close(got[0].cpu(), want, rtol=0, atol=2e-5)


@pytest.mark.parametrize("case", range(5))
def test_duke_transform_matches_reference_golden(case):
    """Against the output of the reference's own transform classes (tests/golden/duke_transform_ref.npz, see
    tests/golden/make_transform_golden.py), not the oracle: same tolerance as above."""
    import os
    from new_vit_b200 import duke_transform
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "duke_transform_ref.npz"))
    v, crop = g[f"in{case}"], tuple(int(c) for c in g[f"crop{case}"])
    got = duke_transform(torch.from_numpy(v).cuda(), image_crop=crop)
    torch.testing.assert_close(got[0].cpu(), torch.from_numpy(g[f"out{case}"]), rtol=0, atol=2e-5)


def test_duke_transform_batch_items_are_independent_and_bit_identical():
    from new_vit_b200 import duke_transform
    vols = np.stack([_volume((240, 200, 30), seed=10 + i, kind=k) for i, k in enumerate(["gamma", "ct", "ties"])])
    x = torch.from_numpy(vols).cuda()
    both = duke_transform(x)
    for i in range(3):
        assert torch.equal(both[i], duke_transform(x[i])[0])
    assert torch.equal(duke_transform(x[:, None]), both)         # [items, 1, W, H, D] spelling


def test_duke_transform_properties_full_size():
    """Size-independent properties at the DUKE working size: Flip(1) is the only place the H orientation enters; the masked
    voxels end up with mean 0 / std 1; clamping bounds the output."""
    from new_vit_b200 import duke_transform
    v = torch.from_numpy(_volume((448, 448, 40), seed=99)).cuda()
    y, stats = duke_transform(v, return_stats=True)
    assert torch.equal(y, duke_transform(torch.flip(v, dims=[1]), flip=False))
    mn, mx, lo, hi, mean, std, cnt, status = stats[0].tolist()
    assert status == 0 and cnt > 0.99 * y.numel()
    inner = (y > y.min()) & (y < y.max())                      # strictly inside the clamp
    z = y.double()
    assert abs(float(z.mean())) < 2e-3 and abs(float(z.std()) - 1.0) < 2e-3
    assert float(y.max()) == pytest.approx((hi - mean) / std, rel=1e-5)
    assert float(y.min()) == pytest.approx((lo - mean) / std, rel=1e-5, abs=1e-6)
    assert 0.98 < float(inner.double().mean()) <= 0.991          # 0.5 % clipped at either end


def test_duke_transform_feeds_the_forward():
    """The transform's output is the model's `source`: [B, 1, 32, 224, 224] fp32, accepted by forward()."""
    from new_vit_b200 import DinoV2ClassifierSlice, duke_transform, synth
    sd = synth.make_state_dict("s", out_ch=2, seed=0, variant="init")
    model = DinoV2ClassifierSlice(1, 2, pretrained=False, precision="bf16").cuda().eval()
    model.load_state_dict(sd)
    x = torch.from_numpy(np.stack([_volume((256, 256, 36), seed=s) for s in (1, 2)])).cuda()
    src = duke_transform(x)
    logits = model(src)
    assert logits.shape == (2, 2) and torch.isfinite(logits).all()


def test_duke_transform_errors():
    from new_vit_b200 import duke_transform
    from new_vit_b200._cabi import MSTError
    with pytest.raises(MSTError):
        duke_transform(torch.zeros(8, 8, 8))                                   # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        duke_transform(torch.ones(16, 16, 8, device="cuda"), image_crop=(16, 16, 8))      # constant: empty mask
    x = torch.zeros(16, 16, 8, device="cuda")
    x[0, 0, 0], x[1, 1, 1] = -1.0, 2.0
    with pytest.raises(RuntimeError):
        duke_transform(x, image_crop=(16, 16, 8))                              # masked values all equal: std 0
    with pytest.raises(ValueError):
        duke_transform(torch.zeros(2, 2, 8, 8, 8, device="cuda"))              # two channels


def test_raw_int16_volumes_are_widened_on_the_device():
    """The raw-data route: int16 / uint16 scanner voxels cross PCIe 2 bytes wide and are widened by the library; bit-identical to
    handing over the same values as fp32."""
    import torch
    from new_vit_b200 import duke_transform
    g = torch.Generator().manual_seed(5)
    raw = torch.randint(-200, 3000, (3, 64, 60, 36), generator=g, dtype=torch.int16)
    want = duke_transform(raw.float().cuda(), (56, 56, 32))
    got = duke_transform(raw.cuda(), (56, 56, 32))
    assert torch.equal(got, want)
