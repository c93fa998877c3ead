import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from dsf_b200 import make_synthetic_mano, sample_fit_inputs
from dsf_b200.fit import FitStep
from dsf_b200.mano_layer import MANO_SMPL
from oracle import mano_oracle as mo
from test_gpu_parity import _oracle_render, _inputs
m = make_synthetic_mano(0)
layer = MANO_SMPL(m, "nyu"); c32 = mo.ManoConstants(m)
B, R = 8, 128
inp = _inputs(B, seed=44)
tgt_inp = dict(inp); tgt_inp["params"] = inp["params_target"]
_, tgt_ref, *_ = _oracle_render(c32, tgt_inp, "direct")
step = FitStep(layer, B, R, use_graph=False)
step.set_inputs(inp["params"].cuda(), inp["center3d"].cuda(), inp["cube"].cuda(), tgt_ref.detach().cuda())
step.step(); torch.cuda.synchronize()
PER = 6708
ws = step.ws[0][:B * PER].view(B, PER).double().cpu()
GVP = ws[:, 2996:2996 + 2336]
GX = ws[:, 5524:5524 + 8 * 148].view(B, 8, 148).sum(1)
X = ws[:, 0:148]; VP = ws[:, 148:148 + 2336]
D = torch.zeros(148, 2336, dtype=torch.float64)
D[:10, :2334] = layer.shapedirs.double().cpu(); D[10:145, :2334] = layer.posedirs.double().cpu()
ref = GVP @ D.T
err = (GX - ref).abs().amax(1) / ref.abs().amax(1)
print("bwd gemm rel err per hand", err)
print("bwd cancellation |sum|/sum|.|", (ref.abs().amax(1) / (GVP.abs() @ D.abs().T).amax(1)))
vt = torch.zeros(2336, dtype=torch.float64); vt[:2334] = layer.v_template.double().cpu().flatten()
reff = X @ D + vt
print("fwd gemm rel err", ((VP - reff).abs().amax(1) / reff.abs().amax(1)))
# gradient comparison per parameter group
p, img_ref, p2f_ref, *_ = _oracle_render(c32, inp, "direct")
mask = tgt_ref.detach().lt(0.99) | img_ref.lt(0.99)
per = (torch.abs(tgt_ref.detach() - img_ref) * mask).sum((-1, -2)) / (mask.float().sum((-1, -2)) + 1e-8)
(g_ref,) = torch.autograd.grad(per.mean() * 0.1, p)
g = step.g_params.cpu()
d = (g - g_ref).abs()
sc = g_ref.abs().amax(1, keepdim=True)
print("same p2f", (step.p2f.cpu() == p2f_ref).flatten(1).all(1))
for name, sl in (("quat", slice(0, 3)), ("theta", slice(3, 48)), ("beta", slice(48, 58)), ("cam", slice(58, 62))):
    print(name, (d[:, sl] / sc).amax(1))
print("gmax", sc.flatten())
