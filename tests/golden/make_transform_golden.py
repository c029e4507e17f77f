"""Generates tests/golden/duke_transform_ref.npz by running the reference's OWN input-pipeline classes
(oracle/ref_transform_harness.py: augmentations_3d.py executed from /root/reference over torchio stand-ins) on seeded
volumes.  Container only (needs /root/reference); the fixture travels.  Run: python tests/golden/make_transform_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_transform_harness as R  # noqa: E402

# (W0, H0, D0), image_crop (W, H, D), kind -- crop / pad / mixed on every axis, odd and even amounts, ties at the extremes
CASES = [
    ((40, 36, 10), (32, 32, 8), "gamma"),
    ((24, 28, 6), (32, 32, 8), "ct"),
    ((37, 27, 8), (32, 32, 8), "ties"),
    ((32, 32, 8), (32, 32, 8), "gamma"),
    ((45, 20, 13), (36, 28, 12), "ct"),
]


def volume(shape, seed, kind):
    rng = np.random.default_rng(seed)
    if kind == "gamma":
        return rng.gamma(2.0, 100.0, size=shape).astype(np.float32)
    if kind == "ct":
        v = rng.normal(0.0, 300.0, size=shape).astype(np.float32)
        v[rng.random(shape) < 0.3] = -1024.0
        return v
    return np.round(rng.normal(0, 3, size=shape)).astype(np.float32)


def main():
    out = {"n": np.int64(len(CASES))}
    for i, (shape, crop, kind) in enumerate(CASES):
        v = volume(shape, 100 + i, kind)
        y = R.reference_duke_transform(v, image_crop=crop).numpy()
        assert y.shape == (1, crop[2], crop[1], crop[0]), y.shape
        out[f"in{i}"], out[f"crop{i}"], out[f"out{i}"] = v, np.asarray(crop, np.int64), y
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "duke_transform_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
