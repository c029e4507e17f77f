"""CPU checks of oracle/duke_transform_oracle.py (the input-pipeline oracle).  torchio is not in this
image; the reference's OWN transform code (augmentations_3d.py) is run in place over torchio stand-ins by
oracle/ref_transform_harness.py and its outputs are the fixture tests/golden/duke_transform_ref.npz, to which the oracle must be
equal (last two tests).  The rest pins what torchio does around it: the closed form the CUDA gather uses for np.pad's 'minimum'
mode against np.pad itself, the bounds rule, and the statistics contract of the z-normalisation."""
import numpy as np
import pytest
import torch

from oracle import duke_transform_oracle as O


def test_six_bounds_follow_the_reference_rule():
    # augmentations_3d.py:166-175: ini = ceil(n / 2), fin = n - ini
    assert O.six_bounds([0, 1, 2]) == (0, 0, 1, 0, 1, 1)
    assert O.six_bounds([5, 24, 7]) == (3, 2, 12, 12, 4, 3)


def _closed_form_minimum_pad(x, pads):
    """Closed form of np.pad(mode='minimum') used by csrc/prep.cu: a padded voxel whose axes in S are out of range holds the
    minimum over the source with the in-range coordinates fixed and the axes in S free."""
    (w0, w1), (h0, h1), (d0, d1) = pads
    W, H, D = x.shape
    out = np.empty((W + w0 + w1, H + h0 + h1, D + d0 + d1), x.dtype)
    for w in range(out.shape[0]):
        sw = w - w0
        iw = 0 <= sw < W
        for h in range(out.shape[1]):
            sh = h - h0
            ih = 0 <= sh < H
            sub = x[sw if iw else slice(None)]
            sub = sub[..., sh, :] if ih else sub.min(axis=-2)
            if not iw:
                sub = sub.min(axis=0)     # [D]
            for d in range(out.shape[2]):
                sd = d - d0
                out[w, h, d] = sub[sd] if 0 <= sd < D else sub.min()
    return out


@pytest.mark.parametrize("pads", [((2, 1), (0, 0), (0, 0)), ((0, 0), (1, 2), (3, 0)), ((2, 2), (1, 1), (2, 1))])
def test_minimum_pad_closed_form_equals_numpy(pads):
    rng = np.random.default_rng(3)
    x = rng.normal(size=(5, 6, 4)).astype(np.float32)
    assert np.array_equal(_closed_form_minimum_pad(x, pads), np.pad(x, pads, mode="minimum"))


@pytest.mark.parametrize("shape", [(40, 36, 10), (20, 50, 7), (30, 28, 8), (31, 29, 9)])
def test_crop_or_pad_shapes_and_centering(shape):
    rng = np.random.default_rng(5)
    x = rng.normal(size=(1,) + shape).astype(np.float32)
    target = (30, 28, 8)
    y = O.crop_or_pad(x, target)
    assert y.shape == (1,) + target
    # the voxel at source index `crop_ini` (or target index `pad_ini`) is the first one kept on each axis
    src_i = tuple(int(np.ceil(max(s - t, 0) / 2)) for s, t in zip(shape, target))
    dst_i = tuple(int(np.ceil(max(t - s, 0) / 2)) for s, t in zip(shape, target))
    assert y[(0,) + dst_i] == x[(0,) + src_i]
    assert y.min() >= x.min() and y.max() <= x.max()      # 'minimum' padding never invents a value


def test_znorm_statistics_contract():
    g = torch.Generator().manual_seed(2)
    x = torch.rand((1, 24, 20, 8), generator=g) * 300 + torch.randn((1, 24, 20, 8), generator=g).abs() * 50
    y, st = O.znorm_percentile(x, (0.5, 99.5))
    mask = (x > x.min()) & (x < x.max())
    assert st["count"] == int(mask.sum()) == x.numel() - 2
    vals = x[mask]
    want = torch.quantile(vals, torch.tensor([0.005, 0.995]))
    assert st["lo"] == float(want[0]) and st["hi"] == float(want[1])
    assert abs(float(y[mask].mean())) < 1e-5 and abs(float(y[mask].std()) - 1) < 1e-5
    assert float(y.max()) == pytest.approx((st["hi"] - st["mean"]) / st["std"], rel=1e-6)    # clamped before normalising


def test_znorm_raises_like_the_reference():
    with pytest.raises(RuntimeError):        # constant volume: empty mask -> torch.quantile raises (augmentations_3d.py:75)
        O.znorm_percentile(torch.ones(1, 4, 4, 4))
    x = torch.zeros(1, 4, 4, 4)
    x[0, 0, 0, 0], x[0, 1, 1, 1] = -1.0, 2.0   # all masked voxels equal -> std 0 -> RuntimeError (:79-84)
    with pytest.raises(RuntimeError):
        O.znorm_percentile(x)


def test_duke_transform_layout():
    rng = np.random.default_rng(7)
    v = rng.gamma(2.0, 100.0, size=(36, 30, 6)).astype(np.float32)
    src, st = O.duke_transform(v, image_crop=(28, 28, 8))
    assert src.shape == (1, 8, 28, 28) and src.dtype == torch.float32
    # out[0, d, h, w] comes from flipped / cropped / padded in[w, h, d]: crop 8 in W (ini 4), 2 in H (ini 1), pad 2 in D (ini 1)
    raw = (v[4 + 3, 30 - 1 - (1 + 5), 2] - st["mean"]) / st["std"]
    raw = (min(max(v[4 + 3, 30 - 1 - (1 + 5), 2], st["lo"]), st["hi"]) - st["mean"]) / st["std"]
    assert float(src[0, 1 + 2, 5, 3]) == pytest.approx(raw, rel=1e-5, abs=1e-6)


# ---- pinned against the reference's own code (tests/golden/duke_transform_ref.npz, made by make_transform_golden.py) ----
def _golden_cases():
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "duke_transform_ref.npz"))
    return [(g[f"in{i}"], tuple(int(c) for c in g[f"crop{i}"]), torch.from_numpy(g[f"out{i}"])) for i in range(int(g["n"]))]


@pytest.mark.parametrize("case", range(5))
def test_oracle_matches_reference_golden(case):
    """The fixture is the output of the reference's OWN CropOrPad / ZNormalization / ImageOrSubjectToTensor objects
    (augmentations_3d.py run in place over torchio stand-ins, oracle/ref_transform_harness.py).  Same ATen ops on both sides:
    bit-equal in the container that made it; 1e-6 leaves room for another CPU's reduction order in mean / std."""
    v, crop, want = _golden_cases()[case]
    got, _ = O.duke_transform(v, image_crop=crop)
    assert got.shape == want.shape
    torch.testing.assert_close(got, want, rtol=0, atol=1e-6)


def test_reference_code_reproduces_the_fixture():
    """Container only: run the reference's classes again and compare with the committed fixture and with the oracle, bit for bit."""
    from oracle import ref_transform_harness as R
    if not R.reference_available():
        pytest.skip("/root/reference is not present (GPU box)")
    import sys
    for v, crop, want in _golden_cases():
        ref = R.reference_duke_transform(v, image_crop=crop)
        assert torch.equal(ref, want)
        assert torch.equal(O.duke_transform(v, image_crop=crop)[0], ref)
    assert "torchio" not in sys.modules or hasattr(sys.modules["torchio"], "__version__")   # the stand-ins are gone again
    A = R.load_reference_augmentations()
    c = A.CropOrPad((8, 8, 8), random_center=False)
    for n in ([0, 1, 2], [5, 24, 7]):       # the oracle's bounds rule against the reference's own method (:166-175)
        assert tuple(int(b) for b in c._get_six_bounds_parameters(np.array(n))) == O.six_bounds(n)
