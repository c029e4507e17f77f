"""One MST-DINOv2 bf16 forward (config 2 shape by default) inside a cudaProfiler range, for ncu:
   ncu --profile-from-start off ... python profiles/prof_forward.py [--batch 64] [--saliency]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from new_vit_b200 import DinoV2ClassifierSlice, synth

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--saliency", action="store_true")
a = ap.parse_args()
m = DinoV2ClassifierSlice(1, 2, pretrained=False, precision="bf16").cuda().eval()
m.load_state_dict(synth.make_state_dict("s", 2, seed=0))
x = torch.randn(a.batch, 1, 32, 224, 224, device="cuda")
with torch.no_grad():
    for _ in range(2):
        y = m(x, save_attn=a.saliency)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    y = m(x, save_attn=a.saliency)
    if a.saliency:
        m.saliency_volume()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("logits[0]", y[0].tolist())
