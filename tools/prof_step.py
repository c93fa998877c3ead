"""Run a few fitting steps for profiling under ncu: python tools/prof_step.py [B] [steps] [graph]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsf_b200 import make_synthetic_mano, sample_fit_inputs
from dsf_b200.fit import FitStep
from dsf_b200.mano_layer import MANO_SMPL
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
layer = MANO_SMPL(make_synthetic_mano(0), "nyu")
inp = {k: torch.from_numpy(v).cuda() for k, v in sample_fit_inputs(B, seed=1000).items()}
s = FitStep(layer, B, 128, use_graph=False)
s.set_inputs(inp["params"], inp["center3d"], inp["cube"])
s.render_target(inp["params_target"])
for _ in range(steps):
    s.step()
torch.cuda.synchronize()
print("loss", float(s.totals[0]))
