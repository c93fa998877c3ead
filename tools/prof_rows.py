"""ncu target: the fused step with the row-run target decoded in the rasteriser's epilogue."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsf_b200 import make_synthetic_mano, sample_fit_inputs
from dsf_b200.fit import FitStep
from dsf_b200.mano_layer import MANO_SMPL
from dsf_b200.pcl import pack_target_rows
from dsf_b200.synthetic import quantise_depth_mm

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
layer = MANO_SMPL(make_synthetic_mano(0), "nyu")
inp = {k: torch.from_numpy(v).cuda() for k, v in sample_fit_inputs(B, seed=1000).items()}
a = FitStep(layer, B, 128, use_graph=False, chunks=1, keep_pix_to_face=False)
a.set_inputs(inp["params"], inp["center3d"], inp["cube"])
a.render_target(inp["params_target"])
mm = quantise_depth_mm(a.target, a.center3d, a.cube)
a.set_inputs(inp["params"], inp["center3d"], inp["cube"], mm)
packed = pack_target_rows(mm.cpu(), inp["center3d"].cpu(), inp["cube"].cpu())
b = FitStep(layer, B, 128, use_graph=False, chunks=1, keep_pix_to_face=False, fuse_target_rows=True)
b.set_inputs(inp["params"], inp["center3d"], inp["cube"], packed)
for _ in range(3):
    a.step()
    b.step()
torch.cuda.synchronize()
print("loss", float(a.totals[0]), float(b.totals[0]))
