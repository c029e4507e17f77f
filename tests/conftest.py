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


CTOR_KEYS = ("use_bottleneck", "use_slice_pos_emb", "slice_fusion", "enable_linear")


def case_inputs(meta, out_ch=2):
    from new_vit_b200 import synth
    hub = bool(meta.get("hub_layout", False))
    ctor = {k: meta[k] for k in CTOR_KEYS if k in meta}
    sd = synth.make_state_dict(meta["size"], out_ch=out_ch, seed=meta["wseed"], variant=meta["variant"],
                               img_size=meta.get("pos_img", meta["H"]), layerscale=hub, chunked_names=not hub,
                               num_registers=meta.get("num_registers", 0), rope=bool(meta.get("rope", False)),
                               liere=bool(meta.get("liere", False)), **ctor)
    x = synth.make_volume(meta["B"], meta["D"], meta["H"], meta["W"], seed=meta["vseed"])
    mask = synth.make_padding_mask(meta["B"], meta["D"], seed=meta["vseed"]) if meta["masked"] else None
    if meta.get("mask_tail"):
        import torch
        mask = torch.zeros(meta["B"], meta["D"], dtype=torch.bool)
        mask[:, meta["D"] - meta["mask_tail"]:] = True
    return sd, x, mask


def model_kwargs(meta):
    """Constructor arguments of new_vit_b200.DinoV2ClassifierSlice for a golden case."""
    kw = {k: meta[k] for k in CTOR_KEYS if k in meta}
    kw.update(model_size=meta["size"], img_size=meta.get("pos_img", meta["H"]), hub_layout=bool(meta.get("hub_layout", False)),
              use_registers=meta.get("num_registers", 0) > 0)
    if meta.get("rope"):
        kw["rotary_positional_encoding"] = "RoPE"
    if meta.get("liere"):
        kw["rotary_positional_encoding"] = "LiRE"
    return kw


@pytest.fixture(scope="session")
def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


@pytest.fixture(autouse=True)
def _seed_everything():
    """Every test starts from the same RNG state (some draw inputs from the global generator): a test either passes or fails,
    it does not do so one run in ten."""
    import torch
    torch.manual_seed(1234)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(1234)
    yield

