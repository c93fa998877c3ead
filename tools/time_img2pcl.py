"""CUDA-event timing of dsf_img2pcl on 1024 synthetic crops (20 % foreground)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsf_b200.pcl import Img2pcl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
img = torch.where(torch.rand(B, 1, 128, 128, device="cuda") < 0.2, torch.rand(B, 1, 128, 128, device="cuda") * 1.6 - 0.8,
                  torch.ones(B, 1, 128, 128, device="cuda"))
c = torch.tensor([[0., 0., 800.]], device="cuda").repeat(B, 1)
q = torch.full((B, 3), 250., device="cuda")
M = torch.eye(3, device="cuda").repeat(B, 1, 1)
M[:, 0, 0] = M[:, 1, 1] = 0.4
M[:, 0, 2], M[:, 1, 2] = -60, -30
for S in (2048, 0):
    for _ in range(3):
        Img2pcl(img, 128, c, M, q, S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        Img2pcl(img, 128, c, M, q, S)
    e1.record()
    torch.cuda.synchronize()
    print("img2pcl hands", B, "sample_num", S, "ms", e0.elapsed_time(e1) / 50)
