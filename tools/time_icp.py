"""CUDA-event timing of the point-face kernels (config C4): forward, backward, with / without the
spatial order.  python tools/time_icp.py [batch] [ordered_cloud 0|1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsf_b200 import _lib as L, make_synthetic_mano, sample_fit_inputs
from dsf_b200.mano_layer import MANO_SMPL

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ordered = len(sys.argv) > 2 and sys.argv[2] == "1"
P = 2048
lib = L.lib()
layer = MANO_SMPL(make_synthetic_mano(0), "nyu")
p = torch.from_numpy(sample_fit_inputs(B, seed=5)["params"]).cuda()
v, _ = layer.get_mano_vertices(p[:, :3], p[:, 3:48], p[:, 48:58], p[:, 58:], global_scale=1 / 125)
v = v.detach().contiguous()
gen = torch.Generator(device="cuda").manual_seed(0)
idx = torch.randint(0, 778, (B, P), device="cuda", generator=gen)
if ordered:
    idx = idx.sort(1)[0]
pcl = (torch.gather(v, 1, idx[..., None].expand(-1, -1, 3)) + 0.05 * torch.randn(B, P, 3, device="cuda", generator=gen)).contiguous()
faces = layer.faces_int
d = torch.empty(B, P, device="cuda")
i = torch.empty(B, P, dtype=torch.int32, device="cuda")
order = torch.empty(B * (P + faces.shape[0]), dtype=torch.int32, device="cuda")
g = torch.ones(B, P, device="cuda")
gp, gv = torch.empty_like(pcl), torch.empty_like(v)


def timeit(fn, n=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


fwd = lambda o: L.check(lib.dsf_point_face_forward(B, P, v.shape[1], faces.shape[0], pcl.data_ptr(), v.data_ptr(), faces.data_ptr(),
                                                   d.data_ptr(), i.data_ptr(), o, L.stream_ptr()))
bwd = lambda: L.check(lib.dsf_point_face_backward(B, P, v.shape[1], faces.shape[0], pcl.data_ptr(), v.data_ptr(), faces.data_ptr(),
                                                  i.data_ptr(), g.data_ptr(), gp.data_ptr(), gv.data_ptr(), L.stream_ptr()))
print("B", B, "fwd sorted", timeit(lambda: fwd(order.data_ptr())), "fwd unsorted", timeit(lambda: fwd(None)), "bwd", timeit(bwd))
