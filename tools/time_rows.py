"""Resident step time with the fp32 target plane vs the row-run target decoded in the rasteriser's epilogue.
usage: python tools/time_rows.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsf_b200 import make_synthetic_mano, sample_fit_inputs
from dsf_b200.fit import FitStep
from dsf_b200.mano_layer import MANO_SMPL
from dsf_b200.pcl import pack_target_rows
from dsf_b200.synthetic import quantise_depth_mm

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else (2 if B >= 2048 else 1)
layer = MANO_SMPL(make_synthetic_mano(0), "nyu")
inp = {k: torch.from_numpy(v).cuda() for k, v in sample_fit_inputs(B, seed=1000).items()}
a = FitStep(layer, B, 128, use_graph=True, chunks=chunks, keep_pix_to_face=False)
a.set_inputs(inp["params"], inp["center3d"], inp["cube"])
a.render_target(inp["params_target"])
mm = quantise_depth_mm(a.target, a.center3d, a.cube)
a.set_inputs(inp["params"], inp["center3d"], inp["cube"], mm)
packed = pack_target_rows(mm.cpu(), inp["center3d"].cpu(), inp["cube"].cpu())
b = FitStep(layer, B, 128, use_graph=True, chunks=chunks, keep_pix_to_face=False, fuse_target_rows=True)
b.set_inputs(inp["params"], inp["center3d"], inp["cube"], packed)
for name, s in (("fp32 plane", a), ("row-run fused", b), ("fp32 plane", a), ("row-run fused", b)):
    for _ in range(5):
        s.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        s.step()
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B} {name:14s} {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us/step")
assert torch.equal(a.g_params, b.g_params)
