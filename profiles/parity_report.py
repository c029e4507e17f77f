"""Print the parity margins of the CUDA path against the reference-generated golden vectors (both precisions)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from conftest import case_inputs, load_golden
from new_vit_b200 import DinoV2ClassifierSlice

for name in ["s_init_b2", "s_peaky_mask_b2", "b_peaky_252_mask_b2", "s_hub_layerscale_b1"]:
    meta, g = load_golden(name)
    sd, x, mask = case_inputs(meta)
    for prec in ("fp32", "bf16"):
        m = DinoV2ClassifierSlice(1, 2, pretrained=False, precision=prec, img_size=meta["H"], model_size=meta["size"],
                                  hub_layout=meta.get("hub_layout", False)).cuda().eval()
        m.load_state_dict(sd)
        with torch.no_grad():
            y = m(x, save_attn=True, src_key_padding_mask=mask).cpu()
            maps = m.get_attention_maps().cpu()
        err = (y - g["logits"]).abs().max().item()
        cos = F.cosine_similarity(maps.flatten().double(), g["attn_maps"].flatten().double(), dim=0).item()
        same = bool(torch.equal(maps.mean(1).reshape(meta["B"], -1).argmax(-1), g["attn_maps"].mean(1).reshape(meta["B"], -1).argmax(-1)))
        print(f"{name:24s} {prec}: max|dlogit| {err:.2e}  map cosine {cos:.6f}  argmax voxel equal {same}")
