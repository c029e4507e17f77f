import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    import ast
    import numpy as np
    import torch
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = ast.literal_eval(str(z["meta"]))
    return meta, {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}


def case_inputs(meta, out_ch=2):
    from new_vit_b200 import synth
    hub = bool(meta.get("hub_layout", False))
    sd = synth.make_state_dict(meta["size"], out_ch=out_ch, seed=meta["wseed"], variant=meta["variant"],
                               img_size=meta["H"], layerscale=hub, chunked_names=not hub)
    x = synth.make_volume(meta["B"], meta["D"], meta["H"], meta["W"], seed=meta["vseed"])
    mask = synth.make_padding_mask(meta["B"], meta["D"], seed=meta["vseed"]) if meta["masked"] else None
    return sd, x, mask


@pytest.fixture(scope="session")
def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))
