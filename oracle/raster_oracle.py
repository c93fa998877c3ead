"""TEST INFRASTRUCTURE ONLY - ctypes bindings + torch glue for oracle/raster_oracle.c.

View set-up restates, in float32 torch ops, Render.points3DToImg / comToBounds / Offset2Trans
(mano_layer.py:1318-1324, :1133-1169) and the pytorch3d-0.4.0 screen-space camera -> NDC
convention.  Two sampling modes (Appendix A):
  "direct"  - an R x R raster with crop-space intrinsics, samples at crop pixel centres;
  "literal" - the reference's S x S raster -> resize -> crop chain, evaluated only at the one
              raster pixel each crop pixel ends up reading.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import build as _build
from . import mano_oracle as mo

_lib = None
VIEW_STRIDE = 8   # fxn, fyn, pxn, pyn, zc, zhalf, bg, pad

# RasterizationSettings.perspective_correct as the reference gets it: mano_layer.py:946-950 builds
# RasterizationSettings(image_size, blur_radius=0.0, faces_per_pixel=1) and nothing else, and in
# pytorch3d 0.4.0 (README.md:41) the constructor default is `perspective_correct: bool = False`, passed on
# unchanged by MeshRasterizer.forward (the inference "True for perspective cameras" arrived in 0.5.0 together
# with Optional[bool] = None).  Recalled, not vendored - so both settings stay testable; every function below
# takes perspective_correct=None meaning this default.
PERSPECTIVE_CORRECT = False


def _pc(flag):
    return PERSPECTIVE_CORRECT if flag is None else bool(flag)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
    return _lib


def _p(t, ty=ctypes.c_float):
    if t is None:
        return None
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ty))


def pix_to_ndc(n: int) -> torch.Tensor:
    """pytorch3d PixToNdc of the flipped index: sample i sits at -1 + (2 (n-1-i) + 1)/n."""
    i = torch.arange(n, dtype=torch.float32)
    return -1.0 + (2.0 * (n - 1 - i) + 1.0) / n


def make_view(mode, center3d, cube, intr, W, H, crop, S=None):
    """-> view (B,8) float32 [fxn,fyn,pxn,pyn,zc,zhalf,bg,0], xs (B,crop), ys (B,crop), M (B,3,3)."""
    center3d = center3d.float()
    cube = cube.float()
    B = center3d.shape[0]
    fx, fy, px, py = intr
    c2 = mo.points3d_to_img(center3d, intr)
    x0, x1, y0, y1 = mo.com_to_bounds(c2, cube, intr)
    M = mo.offset_to_trans(x0, x1, y0, y1, crop)
    view = torch.zeros(B, VIEW_STRIDE)
    zc = center3d[:, 2]
    zh = cube[:, 2] / 2.
    view[:, 4] = zc
    view[:, 5] = zh
    view[:, 6] = ((zc + zh) - zc) / zh            # what normalize_img (:1289-1299) gives background
    if mode in ("direct", "direct_aligned"):
        s = M[:, 0, 0]
        fxc, fyc = s * fx, s * fy
        pxc, pyc = s * px + M[:, 0, 2], s * py + M[:, 1, 2]
        if mode == "direct_aligned":        # sample i at crop coordinate i (M's convention), not i + 1/2
            pxc, pyc = pxc + 0.5, pyc + 0.5
        half = crop / 2.
        view[:, 0] = fxc / half
        view[:, 1] = fyc / half
        view[:, 2] = -(pxc - half) / half
        view[:, 3] = -(pyc - half) / half
        xs = pix_to_ndc(crop)[None].repeat(B, 1)
        ys = xs.clone()
    elif mode == "literal":
        S = S or max(W, H)
        view[:, 0] = fx / (W / 2.)
        view[:, 1] = fy / (H / 2.)
        view[:, 2] = -(px - W / 2.) / (W / 2.)
        view[:, 3] = -(py - H / 2.) / (H / 2.)
        xi, yi, _, _ = mo.literal_sample_maps(M, W, H, S, crop)
        ndc = torch.cat([pix_to_ndc(S), torch.tensor([float("nan")])])
        xs = ndc[xi]          # index -1 -> NaN = reads zero padding
        ys = ndc[yi]
    else:
        raise ValueError(mode)
    return view.contiguous(), xs.contiguous(), ys.contiguous(), M


def normalize_depth(zbuf, view):
    """mano_layer.py:1084-1085 (bg -1 -> 0) then normalize_img (:1289-1299)."""
    z = torch.where(zbuf <= 0, torch.zeros_like(zbuf), zbuf)
    return mo.normalize_img(z[:, None], view[:, 4], view[:, 5] * 2)[:, 0]


def render(verts, faces, view, xs, ys, perspective_correct=None, eps=1e-8, zcull_mode=0, want_bary=False):
    """verts (B,V,3) camera-space mm, faces (F,3) int32 -> pix_to_face (B,R,R) i32, zbuf, [bary], vndc."""
    verts = verts.detach().float().contiguous()
    faces = faces.int().contiguous()
    B, V, _ = verts.shape
    R = xs.shape[1]
    p2f = torch.empty(B, R, R, dtype=torch.int32)
    zbuf = torch.empty(B, R, R)
    bary = torch.empty(B, R, R, 3) if want_bary else None
    vndc = torch.empty(B, V, 3)
    lib().orc_batch_render_f32(
        _p(verts), B, V, _p(faces, ctypes.c_int), faces.shape[0], _p(view), _p(xs), _p(ys), R,
        int(_pc(perspective_correct)), ctypes.c_float(eps), zcull_mode, _p(p2f, ctypes.c_int), _p(zbuf),
        _p(bary), _p(vndc))
    return p2f, zbuf, bary, vndc


def render_f64(verts, faces, view, xs, ys, perspective_correct=None, eps=1e-8, zcull_mode=0):
    """Same rasteriser evaluated in float64 from the same float32 inputs: pixels where it
    disagrees with the float32 build are rounding-ambiguous (tie class T2)."""
    L = lib()
    verts = verts.detach().float().contiguous()
    faces = faces.int().contiguous()
    B, V, _ = verts.shape
    R = xs.shape[1]
    p2f = torch.empty(B, R, R, dtype=torch.int32)
    zbuf = torch.empty(B, R, R, dtype=torch.float64)
    xs64, ys64 = xs.double().contiguous(), ys.double().contiguous()
    for b in range(B):
        vn = torch.empty(V, 3, dtype=torch.float64)
        L.orc_project_f64(_p(verts[b]), V, ctypes.c_float(view[b, 0]), ctypes.c_float(view[b, 1]),
                          ctypes.c_float(view[b, 2]), ctypes.c_float(view[b, 3]), _p(vn, ctypes.c_double))
        L.orc_rasterize_f64(_p(vn, ctypes.c_double), _p(faces, ctypes.c_int), faces.shape[0],
                            _p(xs64[b], ctypes.c_double), R, _p(ys64[b], ctypes.c_double), R,
                            int(_pc(perspective_correct)), ctypes.c_double(eps), zcull_mode,
                            _p(p2f[b], ctypes.c_int), _p(zbuf[b], ctypes.c_double), None, None)
    return p2f, zbuf


def render_backward(verts, faces, view, xs, ys, p2f, grad_zbuf, vndc, perspective_correct=None, eps=1e-8):
    verts = verts.detach().float().contiguous()
    faces = faces.int().contiguous()
    B, V, _ = verts.shape
    R = xs.shape[1]
    gv = torch.empty(B, V, 3)
    scratch = torch.empty(B, V, 3)
    lib().orc_batch_render_backward_f32(
        _p(verts), B, V, _p(faces, ctypes.c_int), _p(view), _p(xs), _p(ys), R, _p(p2f, ctypes.c_int),
        _p(grad_zbuf.float().contiguous()), int(_pc(perspective_correct)), ctypes.c_float(eps), _p(vndc),
        _p(scratch), _p(gv))
    return gv


def render_backward_f64(verts, faces, view, xs, ys, p2f, grad_zbuf, perspective_correct=None, eps=1e-8, vndc=None):
    """The same backward evaluated in float64 - the yardstick for gradient tolerances: the float32 per-pixel
    chain rule (pytorch3d's and its restatement alike) loses up to ~1e-3 to cancellation on sliver faces, more
    in the literal 640-pixel raster where float32 NDC coordinates resolve only 2e-5 of a pixel.  With vndc (the
    float32 projected vertices the forward really used, as returned by render()) the gradient is taken at exactly
    the point the forward evaluated; without it the vertices are projected in float64.
    -> grad wrt the camera-space vertices (B,V,3) float64."""
    L = lib()
    verts = verts.detach().float().contiguous()
    faces = faces.int().contiguous()
    B, V, _ = verts.shape
    R = xs.shape[1]
    cd, ci, cf = ctypes.c_double, ctypes.c_int, ctypes.c_float
    out = torch.zeros(B, V, 3, dtype=torch.float64)
    xs64, ys64 = xs.double().contiguous(), ys.double().contiguous()
    gz = grad_zbuf.double().contiguous()
    p2f = p2f.int().contiguous()
    for b in range(B):
        v4 = [cf(float(view[b, i])) for i in range(4)]
        if vndc is None:
            vn = torch.empty(V, 3, dtype=torch.float64)
            L.orc_project_f64(_p(verts[b]), V, *v4, _p(vn, cd))
        else:
            vn = vndc[b].double().contiguous()
        gvn = torch.zeros(V, 3, dtype=torch.float64)
        L.orc_rasterize_backward_f64(_p(vn, cd), _p(faces, ci), _p(xs64[b], cd), R, _p(ys64[b], cd), R,
                                     _p(p2f[b], ci), _p(gz[b], cd), None, int(_pc(perspective_correct)), cd(eps),
                                     _p(gvn, cd))
        L.orc_project_backward_f64(_p(verts[b]), V, *v4, _p(gvn, cd), _p(out[b], cd))
    return out


class RasterDepth(torch.autograd.Function):
    """verts (B,V,3) -> raw zbuf (B,R,R) with -1 background; gradient as pytorch3d's
    rasterize_meshes_backward feeds it (zbuf only, mano_layer.py:1084).  The forward is the float32
    restatement (bit-level yardstick); the backward is evaluated in float64 at the float32 forward's own
    projected vertices (render_backward_f64), because the float32 per-pixel chain rule carries up to ~1e-3 of
    cancellation noise - more than the 1e-4 the parity tests ask for.  render_backward() keeps the float32
    restatement; tests/test_raster_oracle_cpu.py bounds its distance from this one."""

    @staticmethod
    def forward(ctx, verts, faces, view, xs, ys, perspective_correct=None):
        p2f, zbuf, _, vndc = render(verts, faces, view, xs, ys, perspective_correct)
        ctx.pc = perspective_correct
        ctx.save_for_backward(verts, faces, view, xs, ys, p2f, vndc)
        ctx.mark_non_differentiable(p2f)
        return zbuf, p2f

    @staticmethod
    def backward(ctx, g_zbuf, _g):
        verts, faces, view, xs, ys, p2f, vndc = ctx.saved_tensors
        g = render_backward_f64(verts, faces, view, xs, ys, p2f, g_zbuf, ctx.pc, vndc=vndc)
        return g.to(verts.dtype), None, None, None, None, None


def point_face(points, verts, faces, eps=1e-8):
    """points (B,P,3), verts (B,V,3), faces (F,3) -> dists (B,P), idxs (B,P) int32 (face in its own mesh)."""
    points = points.detach().float().contiguous()
    verts = verts.detach().float().contiguous()
    faces = faces.int().contiguous()
    B, P, _ = points.shape
    d = torch.empty(B, P)
    idx = torch.empty(B, P, dtype=torch.int32)
    lib().orc_batch_point_face_f32(_p(points), B, P, _p(verts), verts.shape[1], _p(faces, ctypes.c_int),
                                   faces.shape[0], ctypes.c_float(eps), _p(d), _p(idx, ctypes.c_int))
    return d, idx


def point_face_backward(points, verts, faces, idxs, grad_dists, eps=1e-8, double=False):
    points = points.detach().float().contiguous()
    verts = verts.detach().float().contiguous()
    faces = faces.int().contiguous()
    B, P, _ = points.shape
    dt = torch.float64 if double else torch.float32
    ct = ctypes.c_double if double else ctypes.c_float
    fn = lib().orc_point_face_backward_f64 if double else lib().orc_point_face_backward_f32
    gp = torch.zeros(B, P, 3, dtype=dt)
    gv = torch.zeros(B, verts.shape[1], 3, dtype=dt)
    gd = grad_dists.to(dt).contiguous()
    for b in range(B):
        fn(_p(points[b]), P, _p(verts[b]), _p(faces, ctypes.c_int), _p(idxs[b].contiguous(), ctypes.c_int),
           _p(gd[b], ct), ct(eps), _p(gp[b], ct), _p(gv[b], ct))
    return gp, gv


class PointFaceDistance(torch.autograd.Function):
    """meshLoss.py:21-70 with shared topology."""

    @staticmethod
    def forward(ctx, points, verts, faces):
        d, idx = point_face(points, verts, faces)
        ctx.save_for_backward(points, verts, faces, idx)
        return d

    @staticmethod
    def backward(ctx, g):
        points, verts, faces, idx = ctx.saved_tensors
        gp, gv = point_face_backward(points, verts, faces, idx, g)
        return gp, gv, None
