"""Device time of the input pipeline (csrc/prep.cu, `duke_transform`) per batch of DUKE-like volumes, CUDA events on the
launching stream, against its algorithmic bytes (read the kept source region once as fp32 + write the [D,H,W] fp32 volume
once): python profiles/prep_timing.py [items] [W0 H0 D0]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from new_vit_b200 import duke_transform  # noqa: E402

items = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W0, H0, D0 = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (256, 256, 40)
crop = (224, 224, 32)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand((items, W0, H0, D0), device="cuda", generator=g).mul_(800.0)
for _ in range(3):
    y = duke_transform(x, image_crop=crop, check=False)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
ev[0].record()
for i in range(10):
    y = duke_transform(x, image_crop=crop, check=False)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))[5]
kept = min(W0, crop[0]) * min(H0, crop[1]) * min(D0, crop[2])
alg = items * (kept + crop[0] * crop[1] * crop[2]) * 4
print(json.dumps({"kernel": "prepare_volume (13-20 launches)", "items": items, "src": [W0, H0, D0], "crop": list(crop),
                  "ms_per_batch": ms, "volumes_per_s": items / ms * 1e3, "algorithmic_bytes": alg,
                  "achieved_gbs": alg / ms / 1e6, "peak_gbs": 6530.3, "frac": alg / ms / 1e6 / 6530.3}))
