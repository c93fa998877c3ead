"""Drop-in API path (Render.render -> m2d_loss -> backward) for ncu launch lists / timing:
python tools/prof_render_api.py [batch] [mode]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsf_b200 import make_synthetic_mano, sample_fit_inputs
from dsf_b200.mano_layer import Render
from dsf_b200.render_loss import m2d_loss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mode = sys.argv[2] if len(sys.argv) > 2 else "literal"
dev = torch.device("cuda")
rnd = Render(make_synthetic_mano(0), "nyu", (588.03, 587.07, 320.0, 240.0), (640, 480), (128, 128), mode=mode)
i = {k: torch.from_numpy(v).to(dev) for k, v in sample_fit_inputs(B, seed=9).items()}
with torch.no_grad():
    tgt = rnd.render(i["params_target"], i["center3d"], i["cube"])[0].clone()
p = i["params"].clone().requires_grad_(True)


def step():
    img, juvd, jxyz, mesh = rnd.render(p, i["center3d"], i["cube"])
    (m2d_loss(tgt, img) + 1e-3 * juvd.sum()).backward()
    p.grad = None


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    step()
e1.record()
torch.cuda.synchronize()
print("render api", mode, "B", B, "ms/step", e0.elapsed_time(e1) / 20)
