import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from dsf_b200 import make_synthetic_mano
from dsf_b200.fit import FitStep
from dsf_b200.mano_layer import MANO_SMPL
from oracle import mano_oracle as mo, raster_oracle as ro
from test_gpu_parity import _oracle_render, _inputs, NYU
m = make_synthetic_mano(0)
layer = MANO_SMPL(m, "nyu"); c32 = mo.ManoConstants(m); c64 = mo.ManoConstants(m, torch.float64)
B, R = 8, 128
inp = _inputs(B, seed=44)
tgt_inp = dict(inp); tgt_inp["params"] = inp["params_target"]
_, tgt_ref, *_ = _oracle_render(c32, tgt_inp, "direct")
tgt_ref = tgt_ref.detach()
step = FitStep(layer, B, R, use_graph=False)
step.set_inputs(inp["params"].cuda(), inp["center3d"].cuda(), inp["cube"].cuda(), tgt_ref.cuda())
step.step(); torch.cuda.synchronize()
p, img_ref, p2f_ref, *_ = _oracle_render(c32, inp, "direct")
mask = tgt_ref.lt(0.99) | img_ref.lt(0.99)
per = (torch.abs(tgt_ref - img_ref) * mask).sum((-1, -2)) / (mask.float().sum((-1, -2)) + 1e-8)
(g32,) = torch.autograd.grad(per.mean() * 0.1, p)

# float64 pipeline with the visibility (pix_to_face) and the mask/sign pattern of the f32 render held fixed
L = ro.lib(); cd, ci = ctypes.c_double, ctypes.c_int
view, xs, ys, M = ro.make_view("direct", inp["center3d"], inp["cube"], NYU, 640, 480, R)
faces = c32.faces.int().contiguous()
class Z64(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vn):       # (B,V,3) double
        out = torch.zeros(B, R, R, dtype=torch.float64)
        vn = vn.contiguous()
        for b in range(B):
            pf = p2f_ref[b].contiguous(); x = xs[b].double().contiguous(); y = ys[b].double().contiguous()
            # evaluate depth of the fixed face at each fg pixel = forward of backward's recompute: use full raster then trust same faces
            p_tmp = torch.empty(R, R, dtype=torch.int32); z = torch.empty(R, R, dtype=torch.float64)
            L.orc_rasterize_f64(ro._p(vn[b], cd), ro._p(faces, ci), faces.shape[0], ro._p(x, cd), R, ro._p(y, cd), R, 1, cd(1e-8), 0, ro._p(p_tmp, ci), ro._p(z, cd), None, None)
            out[b] = z
            ctx.mism = getattr(ctx, "mism", 0) + int((p_tmp != pf).sum())
        ctx.save_for_backward(vn)
        return out
    @staticmethod
    def backward(ctx, gz):
        (vn,) = ctx.saved_tensors
        g = torch.zeros_like(vn)
        for b in range(B):
            pf = p2f_ref[b].contiguous(); x = xs[b].double().contiguous(); y = ys[b].double().contiguous()
            gb = torch.zeros(vn.shape[1], 3, dtype=torch.float64)
            L.orc_rasterize_backward_f64(ro._p(vn[b].contiguous(), cd), ro._p(faces, ci), ro._p(x, cd), R, ro._p(y, cd), R, ro._p(pf, ci), ro._p(gz[b].contiguous(), cd), None, 1, cd(1e-8), ro._p(gb, cd))
            g[b] = gb
        print("f64 visibility mismatches vs f32:", ctx.mism)
        return g
p64 = inp["params"].double().requires_grad_(True)
q, t, b_, cam = mo.split_params(p64)
v, j = mo.get_mano_vertices(c64, q, t, b_, cam, global_scale=1 / 125)
vw = v * inp["cube"].double()[:, None] / 2 + inp["center3d"].double()[:, None]
vd = view.double()
xn = (vd[:, 0, None] * -vw[..., 0] + vd[:, 2, None] * vw[..., 2]) / vw[..., 2]
yn = (vd[:, 1, None] * -vw[..., 1] + vd[:, 3, None] * vw[..., 2]) / vw[..., 2]
vn = torch.stack([xn, yn, vw[..., 2]], -1)
z = Z64.apply(vn)
zc, zh = vd[:, 4].view(-1, 1, 1), vd[:, 5].view(-1, 1, 1)
fg = p2f_ref >= 0
zz = torch.where(fg, z, zc + zh)
img64 = (torch.minimum(torch.maximum(zz, zc - zh), zc + zh) - zc) / zh
per64 = (torch.abs(tgt_ref.double() - img64) * mask).sum((-1, -2)) / (mask.double().sum((-1, -2)) + 1e-8)
(g64,) = torch.autograd.grad(per64.mean() * 0.1, p64)
g = step.g_params.cpu().double()
sc = g64.abs().amax(1)
print("GPU  vs f64:", ((g - g64).abs().amax(1) / sc))
print("CPU32 vs f64:", ((g32.double() - g64).abs().amax(1) / sc))
print("GPU  vs CPU32:", ((g - g32.double()).abs().amax(1) / sc))
