"""TEST INFRASTRUCTURE ONLY - the reference's CPU path for one model-fitting step, assembled from
the oracle pieces: torch MANO restatement (mano_layer.py:573-693) -> placement in the camera cube
(:1078) -> pytorch3d-0.4.0 naive CPU rasteriser (C, one OpenMP task per hand) -> background fill +
normalize_img (:1084-1085, :1289-1299) -> inline m2d loss (train_render.py:728-732) -> autograd
backward to the (B,62) parameters.  Used by tests as the end-to-end checker and by bench.py as the
timed CPU baseline / `--impl reference` arm (never by the product)."""
from __future__ import annotations

import torch

from . import mano_oracle as mo
from . import raster_oracle as ro

NYU = (588.03, 587.07, 320.0, 240.0)


def render(consts, params, center3d, cube, mode="direct", crop=128, intr=NYU, sensor=(640, 480),
           perspective_correct=None):
    """-> img (B,crop,crop) normalised depth (differentiable wrt params), pix_to_face, view pack."""
    q, t, b, cam = mo.split_params(params)
    verts, joints = mo.get_mano_vertices(consts, q, t, b, cam, global_scale=1 / 125)
    verts_cam = verts * cube[:, None] / 2 + center3d[:, None]
    view, xs, ys, M = ro.make_view(mode, center3d, cube, intr, sensor[0], sensor[1], crop)
    zbuf, p2f = ro.RasterDepth.apply(verts_cam, consts.faces, view, xs, ys, perspective_correct)
    return ro.normalize_depth(zbuf, view), p2f, (view, xs, ys, M), verts, joints


def m2d_loss(real, synth, weight=0.1, thr=0.99):
    mask = real.lt(thr) | synth.lt(thr)
    diff = torch.abs(real - synth) * mask
    per_hand = diff.sum((-1, -2)) / (mask.float().sum((-1, -2)) + 1e-8)
    return per_hand.mean() * weight, per_hand


def fit_step(consts, params, center3d, cube, target, mode="direct", crop=128, perspective_correct=None):
    """One full step; returns (loss, d loss / d params, img)."""
    p = params.detach().clone().requires_grad_(True)
    img, _, _, _, _ = render(consts, p, center3d, cube, mode, crop, perspective_correct=perspective_correct)
    loss, _ = m2d_loss(target, img)
    (g,) = torch.autograd.grad(loss, p)
    return loss.detach(), g, img.detach()
