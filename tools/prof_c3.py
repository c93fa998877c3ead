"""Config C3 (512 hands x 3 views, 256x256) through the modular autograd API, for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dsf_b200 import make_synthetic_mano, sample_fit_inputs
from dsf_b200.mano_layer import Render, RotationPoints
from dsf_b200.render_loss import m2d_loss

dev = torch.device("cuda")
b3, V3, R3 = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 3, 256
rnd3 = Render(make_synthetic_mano(0), "nyu", (588.03, 587.07, 320.0, 240.0), (640, 480), (R3, R3), mode="direct")
i3 = {k: torch.from_numpy(v).to(dev) for k, v in sample_fit_inputs(b3, seed=9).items()}
rot3 = torch.tensor([[0.0, 0.0, 0.0], [0.0, 2 * np.pi / 3, 0.0], [0.0, -2 * np.pi / 3, 0.0]], device=dev).repeat(b3, 1)
c3v, cube3v = i3["center3d"].repeat_interleave(V3, 0), i3["cube"].repeat_interleave(V3, 0)


def images(params):
    v, j = rnd3.mano_layer.get_mano_vertices(params[:, :3], params[:, 3:48], params[:, 48:58], params[:, 58:], global_scale=1 / 125)
    vw = (v * i3["cube"][:, None] / 2 + i3["center3d"][:, None]).repeat_interleave(V3, 0)
    jw = (j * i3["cube"][:, None] / 2 + i3["center3d"][:, None]).repeat_interleave(V3, 0)
    vr, _ = RotationPoints(vw, jw, c3v, rot3)
    return rnd3._rasterize(vr, c3v, cube3v)[0]


with torch.no_grad():
    tgt = images(i3["params_target"]).clone()
pg = i3["params"].clone().requires_grad_(True)
for it in range(3):
    torch.cuda.nvtx.range_push("c3_step")
    m2d_loss(tgt, images(pg)).backward()
    pg.grad = None
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("done")
