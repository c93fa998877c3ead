"""CUDA-event timing of the fused raster launch with parts of its tail switched off (tuning aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsf_b200 import make_synthetic_mano, sample_fit_inputs, _lib as L
from dsf_b200.fit import FitStep
from dsf_b200.mano_layer import MANO_SMPL

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
R = 128
layer = MANO_SMPL(make_synthetic_mano(0), "nyu")
inp = {k: torch.from_numpy(v).cuda() for k, v in sample_fit_inputs(B, seed=1000).items()}
step = FitStep(layer, B, R, use_graph=False, chunks=1, keep_pix_to_face=False)
step.set_inputs(inp["params"], inp["center3d"], inp["cube"])
step.render_target(inp["params_target"])
step.step()
lib = L.lib()
vcam = (step.verts * step.cube[:, None] / 2 + step.center3d[:, None]).contiguous()
ws = torch.empty(lib.dsf_raster_loss_workspace_floats(B, R), device="cuda")
parts, totals = torch.empty(B, 2, device="cuda"), torch.empty(4, device="cuda")
p2f = torch.empty(B, R, R, dtype=torch.int32, device="cuda")
pt = torch.empty(B * 2 * 2, device="cuda")
s = L.stream_ptr()
h = layer._handle

def fused(flags):
    return lambda: L.check(lib.dsf_raster_loss_grad(h, B, vcam.data_ptr(), step.view.data_ptr(), step.xs.data_ptr(),
        step.ys.data_ptr(), R, step.target.data_ptr(), 0.99, 0.1, B, step.img.data_ptr(), None, parts.data_ptr(),
        totals.data_ptr(), None, ws.data_ptr(), flags, s))

def plain(tgt, pf):
    return lambda: L.check(lib.dsf_raster_forward(h, B, vcam.data_ptr(), step.view.data_ptr(), step.xs.data_ptr(),
        step.ys.data_ptr(), R, step.img.data_ptr(), p2f.data_ptr(), None, None, None,
        step.target.data_ptr() if tgt else None, 0.99, pt.data_ptr() if tgt else None, 0, s))

def t(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for name, fn in [("plain fwd (img+p2f)", plain(False, True)), ("fwd + target/loss parts (+p2f)", plain(True, True)),
                 ("fused: no grad", fused(4096)), ("fused: full", fused(0))]:
    print(f"{name:36s} {t(fn):.4f} ms")
