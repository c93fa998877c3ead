"""Graph-replay step time of FitStep at several batch sizes (warm L2 back to back, and with an L2 flush between steps).
usage: python tools/time_steps.py 128 512 1024 4096"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsf_b200 import make_synthetic_mano, sample_fit_inputs
from dsf_b200.fit import FitStep
from dsf_b200.mano_layer import MANO_SMPL

layer = MANO_SMPL(make_synthetic_mano(0), "nyu")
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for B in [int(a) for a in sys.argv[1:]] or [128, 512, 1024, 4096]:
    inp = {k: torch.from_numpy(v).cuda() for k, v in sample_fit_inputs(B, seed=1000).items()}
    for chunks in (1, 2, 3, 4):
        s = FitStep(layer, B, 128, use_graph=True, chunks=chunks, keep_pix_to_face=False)
        s.set_inputs(inp["params"], inp["center3d"], inp["cube"])
        s.render_target(inp["params_target"])
        for _ in range(5):
            s.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 100
        e0.record()
        for _ in range(n):
            s.step()
        e1.record()
        torch.cuda.synchronize()
        warm = e0.elapsed_time(e1) / n
        tot = 0.0
        for _ in range(20):
            flush.fill_(1.0)
            torch.cuda.synchronize()
            e0.record()
            s.step()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        print(f"B={B:5d} chunks={chunks} launches={s.launches_per_step:2d}  warm {warm*1e3:8.1f} us  cold-L2 {tot/20*1e3:8.1f} us  "
              f"-> {B/warm/1e3:.2f} M fits/s warm")
