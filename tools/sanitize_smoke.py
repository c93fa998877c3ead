"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from dsf_b200 import make_synthetic_mano, sample_fit_inputs
from dsf_b200.fit import FitStep
from dsf_b200.intersection import PartTopology, intersect_counts
from dsf_b200.mano_layer import MANO_SMPL, Render
from dsf_b200.mesh_loss import ICPLoss
from dsf_b200.pcl import Img2pcl, target_from_u16, uvdImg2xyzImg
from dsf_b200.synthetic import quantise_depth_mm

B = 6
layer = MANO_SMPL(make_synthetic_mano(0), "nyu")
inp = {k: torch.from_numpy(v).cuda() for k, v in sample_fit_inputs(B, seed=4).items()}
for R in (128, 100):
    s = FitStep(layer, B, R, use_graph=False)
    s.set_inputs(inp["params"], inp["center3d"], inp["cube"])
    s.render_target(inp["params_target"])
    mm = quantise_depth_mm(s.target, s.center3d, s.cube)
    s.set_inputs(inp["params"], inp["center3d"], inp["cube"], mm)
    s.step()
    torch.cuda.synchronize()
    print("fit step R", R, "loss", float(s.totals[0]))
from dsf_b200.pcl import pack_target_rows
s = FitStep(layer, B, 128, use_graph=False, chunks=2, keep_pix_to_face=False, fuse_target_rows=True)
s.set_inputs(inp["params"], inp["center3d"], inp["cube"])
s.render_target(inp["params_target"])
mm = quantise_depth_mm(s.target, s.center3d, s.cube)
s.set_inputs(inp["params"], inp["center3d"], inp["cube"], pack_target_rows(mm.cpu(), inp["center3d"].cpu(), inp["cube"].cpu()))
s.step()                                                   # row-run target decoded in the raster epilogue, two slices
torch.cuda.synchronize()
print("fit step rows", float(s.totals[0]))
s = FitStep(layer, B, 128, use_graph=False)
s.set_inputs(inp["params"], inp["center3d"], inp["cube"])
s.render_target(inp["params_target"])
s.step()
pcl = Img2pcl(s.img, 128, s.center3d, s.M, s.cube, 2048, seed=3)
pcl64, cnt = Img2pcl(s.img, 64, s.center3d, s.M, s.cube, 0)
xyz, xyzn = uvdImg2xyzImg(s.img, s.center3d, s.M, s.cube)
t = target_from_u16(quantise_depth_mm(s.img, s.center3d, s.cube), s.center3d, s.cube)
r = Render(make_synthetic_mano(0), "nyu", (588.03, 587.07, 320.0, 240.0), (640, 480), mode="literal")
out = r.render(inp["params"], inp["center3d"], inp["cube"])
p = inp["params"]
v, j = layer.get_mano_vertices(p[:, :3], p[:, 3:48], p[:, 48:58], p[:, 58:], global_scale=1 / 125)
jg = j.detach().requires_grad_(True)
layer.calculate_coll(jg, v.detach()).backward()
vg = v.detach().requires_grad_(True)
ICPLoss(vg, pcl[:, :512].contiguous(), layer.faces).mean().backward()      # whole-mesh point-face kernel (P >= 512)
ICPLoss(vg, pcl[:, :300].contiguous(), layer.faces).mean().backward()      # chunked point-face kernel
seg = layer.seg_pcl(j.detach(), j.detach(), v.detach(), pcl[:, :256].contiguous())
v_mm = layer.get_mano_vertices(p[:, :3], p[:, 3:48] * 3, p[:, 48:58], p[:, 58:])[0].detach()
iv = intersect_counts(v_mm[:2].contiguous(), PartTopology.synthetic_hand(), 2.0)
torch.cuda.synchronize()
print("ok", float(pcl.abs().sum()), int(cnt.sum()), float(iv["volume"].sum()), int(seg.sum()))
