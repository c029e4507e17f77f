import os, sys, torch
sys.path.insert(0, os.getcwd())
from new_vit_b200 import DinoV2ClassifierSlice, synth
m = DinoV2ClassifierSlice(1, 2, pretrained=False, precision="bf16").cuda().eval()
m.load_state_dict(synth.make_state_dict("s", 2, seed=0))
x32 = synth.make_volume(64, 32, 224, 224, seed=0).pin_memory()
x = x32.to(torch.bfloat16).pin_memory()   # the bf16 source of bench.py's e2e leg
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
for sched in [(4, 12, 48), (4, 60), (8, 56), (2, 10, 52), (2, 6, 56), (16, 48), (4, 28, 32), (1, 3, 12, 48), (64,), (4, 12, 48)]:
    m.h2d_chunk_volumes = sched
    print(sched, round(run(x), 2))
m.h2d_chunk_volumes = (4, 12, 48)
print("fp32 source (4, 12, 48)", round(run(x32), 2))
