"""crop_hand behind the reference's signature (data/render_loader.py:1209-1227, a method of the data
loader there): removes background and arm from a normalised depth crop by keeping only pixels whose
back-projected 3-D point lies in a box around the (teacher) skeleton.  One CUDA kernel
(dsf_crop_hand); the gradient passes through the kept pixels."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


class _CropHand(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, joint, center, M, cube, paras, offsetxy, offsetz, hand_thickness):
        lib = L.lib()
        img_c = L.f32c(img)
        R = img_c.shape[-1]
        if img_c.shape[-2] != R:
            raise ValueError("square images only")
        B = img_c.numel() // (R * R)
        joint, center, M, cube = L.f32c(joint.detach()), L.f32c(center), L.f32c(M), L.f32c(cube)
        out = torch.empty_like(img_c)
        keep = torch.empty(B, R, R, dtype=torch.uint8, device=img_c.device)
        intr = (C.c_float * 4)(*[float(v) for v in paras])
        L.check(lib.dsf_crop_hand(B, R, img_c.data_ptr(), joint.data_ptr(), joint.shape[1], center.data_ptr(),
                                  cube.data_ptr(), M.data_ptr(), intr, float(offsetxy), float(offsetz),
                                  float(hand_thickness), out.data_ptr(), keep.data_ptr(), L.stream_ptr()))
        ctx.save_for_backward(keep)
        return out

    @staticmethod
    def backward(ctx, g):
        (keep,) = ctx.saved_tensors
        return (g * keep.view_as(g).to(g.dtype),) + (None,) * 8


def crop_hand(img, joint, center, M, cube, paras=(588.03, 587.07, 320.0, 240.0), offsetxy=25, offsetz=20,
              hand_thickness=20):
    """Same arguments as ``loader.crop_hand(img, joint, center, M, cube, ...)`` plus the intrinsics
    (``self.paras`` of the loader)."""
    return _CropHand.apply(img, joint, center, M, cube, paras, offsetxy, offsetz, hand_thickness)
