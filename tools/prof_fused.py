"""ncu target: a few launches of the fused raster kernel (dsf_raster_loss_grad) at bench size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsf_b200 import make_synthetic_mano, sample_fit_inputs, _lib as L
from dsf_b200.fit import FitStep
from dsf_b200.mano_layer import MANO_SMPL

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
layer = MANO_SMPL(make_synthetic_mano(0), "nyu")
inp = {k: torch.from_numpy(v).cuda() for k, v in sample_fit_inputs(B, seed=1000).items()}
step = FitStep(layer, B, 128, use_graph=False, chunks=1, keep_pix_to_face=False)
step.set_inputs(inp["params"], inp["center3d"], inp["cube"])
step.render_target(inp["params_target"])
for _ in range(iters):
    step.step()
torch.cuda.synchronize()
print("loss", float(step.totals[0]))
