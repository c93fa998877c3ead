import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsf_b200 import make_synthetic_mano, sample_fit_inputs, _lib as L
from dsf_b200.mano_layer import MANO_SMPL
layer = MANO_SMPL(make_synthetic_mano(0), "nyu")
B, P = 1024, 2048
inp = {k: torch.from_numpy(v).cuda() for k, v in sample_fit_inputs(B, seed=5).items()}
p = inp["params"]
v, j = layer.get_mano_vertices(p[:, :3], p[:, 3:48], p[:, 48:58], p[:, 58:], global_scale=1 / 125)
v = v.detach().contiguous()
gen = torch.Generator(device="cuda").manual_seed(0)
idx = torch.randint(0, 778, (B, P), device="cuda", generator=gen)
pcl = (torch.gather(v, 1, idx[..., None].expand(-1, -1, 3)) + 0.05 * torch.randn(B, P, 3, device="cuda", generator=gen)).contiguous()
F = layer.faces_int.shape[0]
st = torch.zeros(4, dtype=torch.int64, device="cuda")
d = torch.empty(B, P, device="cuda"); i = torch.empty(B, P, dtype=torch.int32, device="cuda")
o = torch.empty(B * (P + F), dtype=torch.int32, device="cuda")
L.check(L.lib().dsf_point_face_stats(B, P, 779, F, pcl.data_ptr(), v.data_ptr(), layer.faces_int.data_ptr(), d.data_ptr(), i.data_ptr(), o.data_ptr(), st.data_ptr(), L.stream_ptr()))
torch.cuda.synchronize()
c, a, e, g = st.tolist()
n = B * P
print("per point: group tests %.1f, individual culls %.1f, interior evals %.2f, edge evals %.1f (of %d faces)" % (g / n, c / n, a / n, e / n, F))
fo = o[B * P:].view(B, F)[0].long()
faces = layer.faces_int.long()
cen = v[0][faces].mean(1)[fo]           # centroids in sorted order
grp = cen[: (F // 8) * 8].view(-1, 8, 3)
ext = (grp - grp.mean(1, keepdim=True)).norm(dim=-1).max(1)[0]
print("group extent (max centroid offset): mean %.3f max %.3f ; mesh bbox %s ; mean nearest dist %.3f" % (ext.mean(), ext.max(), (v[0].max(0)[0] - v[0].min(0)[0]).tolist(), d.sqrt().mean()))
