"""CUDA-event timing of the intersection-volume kernels (row I1): python tools/time_ivox.py [hands] [pitch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsf_b200 import make_synthetic_mano, sample_fit_inputs
from dsf_b200.intersection import PartTopology, intersect_counts
from dsf_b200.mano_layer import MANO_SMPL

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pitch = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
layer = MANO_SMPL(make_synthetic_mano(0), "nyu")
p = torch.from_numpy(sample_fit_inputs(B, seed=5)["params"]).cuda()
v = layer.get_mano_vertices(p[:, :3], p[:, 3:48] * 3, p[:, 48:58], p[:, 58:])[0].detach().contiguous()
topo = PartTopology.synthetic_hand()
out = intersect_counts(v, topo, pitch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = intersect_counts(v, topo, pitch)
e1.record()
torch.cuda.synchronize()
print("ivox hands", B, "pitch", pitch, "ms", e0.elapsed_time(e1) / 5, "mean volume", float(out["volume"].mean()),
      "status", int(out["status"].abs().sum()))
