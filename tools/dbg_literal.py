import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsf_b200 import make_synthetic_mano, sample_fit_inputs, _lib as L
from dsf_b200.mano_layer import MANO_SMPL
from oracle import mano_oracle as mo, raster_oracle as ro
NYU = (588.03, 587.07, 320.0, 240.0)
model = make_synthetic_mano(0)
layer = MANO_SMPL(model, "nyu")
c32 = mo.ManoConstants(model)
lib = L.lib()
B, R, mode = 48, 128, 1
inp = {k: torch.from_numpy(v) for k, v in sample_fit_inputs(B, seed=300 + R + mode).items()}
p = inp["params"].cuda()
v, j = layer.get_mano_vertices(p[:, :3], p[:, 3:48], p[:, 48:58], p[:, 58:], global_scale=1 / 125)
c3, cube = inp["center3d"].cuda(), inp["cube"].cuda()
vw = (v * cube[:, None] / 2 + c3[:, None]).contiguous().detach()
view = torch.empty(B, L.VIEW_STRIDE, device="cuda"); xs = torch.empty(B, R, device="cuda"); ys = torch.empty(B, R, device="cuda"); M = torch.empty(B, 3, 3, device="cuda")
intr = (C.c_float * 4)(*NYU)
L.check(lib.dsf_view_setup(mode, B, c3.data_ptr(), cube.data_ptr(), intr, 640, 480, R, None, view.data_ptr(), xs.data_ptr(), ys.data_ptr(), M.data_ptr(), L.stream_ptr()))
s = L.stream_ptr()
tgt = torch.rand(B, R, R, device="cuda")
pt = torch.empty(B * 4, device="cuda")
def run(with_target):
    img = torch.empty(B, R, R, device="cuda"); p2f = torch.empty(B, R, R, dtype=torch.int32, device="cuda")
    L.check(lib.dsf_raster_forward(layer._handle, B, vw.data_ptr(), view.data_ptr(), xs.data_ptr(), ys.data_ptr(), R, img.data_ptr(), p2f.data_ptr(), None, None, None, tgt.data_ptr() if with_target else None, 0.99, pt.data_ptr() if with_target else None, 0, s))
    torch.cuda.synchronize()
    return img, p2f
p_ref, z_ref, _, _ = ro.render(vw.cpu(), c32.faces, view[:, :8].cpu().contiguous(), xs.cpu(), ys.cpu(), perspective_correct=False)
for wt in (False, True, False, True, False, True):
    img, p2f = run(wt)
    d = (p2f.cpu() != p_ref)
    print("with_target", wt, "mismatch vs oracle", int(d.sum()), d.nonzero()[:4].tolist(), p2f.cpu()[d][:4].tolist(), p_ref[d][:4].tolist())
