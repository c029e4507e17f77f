import os, sys, torch
sys.path.insert(0, os.getcwd())
from new_vit_b200 import DinoV2ClassifierSlice, synth
m = DinoV2ClassifierSlice(1, 2, pretrained=False, precision="bf16").cuda().eval()
m.load_state_dict(synth.make_state_dict("s", 2, seed=0))
x = synth.make_volume(64, 32, 224, 224, seed=0).pin_memory()
xd = x.cuda()
def run(src, n=6):
    with torch.no_grad():
        for _ in range(2): m(src).cpu()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): m(src).cpu()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("device-resident", round(run(xd), 2))
for sched in [(4, 12, 16, 32), (4, 12, 48), (2, 6, 24, 32), (4, 20, 40), (8, 24, 32), (6, 26, 32), (4, 12, 24, 24), (3, 9, 20, 32)]:
    m.h2d_chunk_volumes = sched
    print(sched, round(run(x), 2))
