"""Render losses behind the reference's call signatures.

``depth_loss()(real, synth)`` mirrors render_model/render_loss.py:9-28 (non-smooth branch; the
``smooth=True`` branch of the reference reads an undefined attribute and cannot run).
``m2d_loss(real, synth)`` is the inline masked L1 of train_render.py:728-732 / :770-774.
Both run in one CUDA kernel pair (dsf_depth_loss) and return the gradient through autograd.
"""
from __future__ import annotations

import torch

from . import _lib as L


class _DepthLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, real, synth, mode, thr, weight):
        lib = L.lib()
        real = L.f32c(real.detach())
        synth_c = L.f32c(synth)
        if real.shape != synth_c.shape:
            raise ValueError("real and synth must have the same shape")
        R = real.shape[-1]
        if real.shape[-2] != R:
            raise ValueError("square images only")
        B = real.numel() // (R * R)
        dev = real.device
        parts = torch.empty(B, 2, device=dev)
        totals = torch.empty(4, device=dev)
        g = torch.empty_like(synth_c)
        L.check(lib.dsf_depth_loss(mode, B, R, real.data_ptr(), synth_c.data_ptr(), float(thr), float(weight),
                                   parts.data_ptr(), totals.data_ptr(), g.data_ptr(), L.stream_ptr()))
        ctx.save_for_backward(g)
        ctx.mark_non_differentiable(parts)
        return totals[0], parts

    @staticmethod
    def backward(ctx, g_loss, _g_parts):
        (g,) = ctx.saved_tensors
        return None, g * g_loss, None, None, None


class depth_loss(torch.nn.Module):
    """mean |real - synth| over pixels where both images are foreground (< 0.99)."""

    def __init__(self, beta=0.4, smooth=False):
        super().__init__()
        if smooth:
            raise NotImplementedError("the reference's smooth branch is dead code (undefined sample_rate)")
        self.smooth = smooth

    def forward(self, real, synth):
        loss, _ = _DepthLossFunction.apply(real, synth, 1, 0.99, 1.0)
        return loss


def m2d_loss(real, synth, weight=0.1, thr=0.99, return_per_hand=False):
    """train_render.py:728-732: union-mask L1, normalised per hand, batch mean times ``weight``."""
    loss, parts = _DepthLossFunction.apply(real, synth, 0, thr, weight)
    if return_per_hand:
        return loss, parts[:, 0] / (parts[:, 1] + 1e-8)
    return loss
