"""Where one config-5 training step goes (8 volumes, ViT-S, every parameter trainable): host-side phases timed with a device
synchronise after each (weight re-pack, forward, backward, optimizer)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from new_vit_b200 import DinoV2ClassifierSlice, synth
m = DinoV2ClassifierSlice(1, 2, pretrained=False, precision="bf16").cuda()
m.load_state_dict(synth.make_state_dict("s", 2, seed=0))
m.train()
opt = m.configure_optimizers()[0]
x = synth.make_volume(8, 32, 224, 224, seed=1).cuda()
batch = {"source": x, "target": torch.randint(0, 2, (8,), device="cuda")}
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    if it == 3: torch.cuda.cudart().cudaProfilerStart()   # ncu --profile-from-start off: one steady-state step
    t0 = sync(); opt.zero_grad(); m.sync_weights(); t1 = sync()
    loss = m.training_step(batch, 0); t2 = sync()
    loss.backward(); t3 = sync()
    opt.step(); t4 = sync()
    if it == 3: torch.cuda.cudart().cudaProfilerStop()
    print(f"iter {it}: zero_grad+repack {1e3*(t1-t0):.2f} ms  forward {1e3*(t2-t1):.2f}  backward {1e3*(t3-t2):.2f}  adamw {1e3*(t4-t3):.2f}  loss {float(loss.detach()):.4f}")
m.profile_begin()
loss = m.training_step(batch, 0); loss.backward()
print({k: (round(v[0], 3), v[1]) for k, v in m.profile_end().items() if v[1]})
