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


class _ChamferNN(torch.autograd.Function):
    """squared distance of every point of x to its nearest point of y and vice versa (dsf_chamfer_forward /
    _backward) - the K = 1 nearest-neighbour core of pytorch3d's chamfer_distance."""

    @staticmethod
    def forward(ctx, x, y):
        lib = L.lib()
        x, y = L.f32c(x), L.f32c(y)
        if x.dim() != 3 or y.dim() != 3 or x.shape[0] != y.shape[0] or x.shape[2] != 3 or y.shape[2] != 3:
            raise ValueError("x (B,P1,3) and y (B,P2,3) expected")
        B, P1, P2 = x.shape[0], x.shape[1], y.shape[1]
        dx = torch.empty(B, P1, device=x.device)
        dy = torch.empty(B, P2, device=x.device)
        ix = torch.empty(B, P1, dtype=torch.int32, device=x.device)
        iy = torch.empty(B, P2, dtype=torch.int32, device=x.device)
        L.check(lib.dsf_chamfer_forward(B, P1, P2, x.data_ptr(), y.data_ptr(), dx.data_ptr(), ix.data_ptr(),
                                        dy.data_ptr(), iy.data_ptr(), L.stream_ptr()))
        ctx.save_for_backward(x, y, ix, iy)
        ctx.mark_non_differentiable(ix, iy)
        return dx, dy, ix, iy

    @staticmethod
    def backward(ctx, g_dx, g_dy, _a, _b):
        lib = L.lib()
        x, y, ix, iy = ctx.saved_tensors
        B, P1, P2 = x.shape[0], x.shape[1], y.shape[1]
        g_dx = torch.zeros(B, P1, device=x.device) if g_dx is None else L.f32c(g_dx)
        g_dy = torch.zeros(B, P2, device=x.device) if g_dy is None else L.f32c(g_dy)
        gx, gy = torch.empty_like(x), torch.empty_like(y)
        L.check(lib.dsf_chamfer_backward(B, P1, P2, x.data_ptr(), y.data_ptr(), ix.data_ptr(), iy.data_ptr(),
                                         g_dx.data_ptr(), g_dy.data_ptr(), gx.data_ptr(), gy.data_ptr(), L.stream_ptr()))
        return gx, gy


def chamfer_distance(x, y):
    """pytorch3d.loss.chamfer_distance(x, y) with its defaults, as render_loss.py:50 calls it: full clouds,
    point_reduction='mean', batch_reduction='mean', no normals -> (loss, None)."""
    dx, dy, _, _ = _ChamferNN.apply(x, y)
    return dx.mean(1).mean() + dy.mean(1).mean(), None


class surface_loss(torch.nn.Module):
    """render_model/render_loss.py:37-52: chamfer distance between the point cloud of the real depth crop and the
    mesh vertices.  Unused by the trainer (train_render.py:16 only imports it), kept importable and working: the
    point cloud comes from the Img2pcl kernel (1024 points, the layout of :53-88; its draw is the kernel's counter
    hash, not torch.multinomial's stream), the chamfer core from dsf_chamfer_*."""

    def __init__(self):
        super().__init__()
        self.img_size = 128
        self.paras = (588.03, 587.07, 320.0, 240.0)
        self.flip = 1

    def Img2pcl(self, img, center, M, cube, verts=None, sample_num=1024, seed=0):
        from .pcl import Img2pcl as _img2pcl

        pcl, count = _img2pcl(img, self.img_size, center, M, cube, sample_num, paras=self.paras,
                              img_size=self.img_size, flip=float(self.flip), seed=seed, return_count=True)
        # the reference gives up on the whole batch when one crop has fewer than two points (:74-76)
        if verts is not None and bool((count < 2).any()):
            return verts
        return pcl

    def forward(self, real, synth, verts, faces, center, M, cube):
        pcl1 = self.Img2pcl(real, center, M, cube, verts)
        loss, _ = chamfer_distance(pcl1, verts)
        return loss
