"""The pytorch3d-shaped lower boundary: ``rasterize_meshes`` / ``Fragments`` for callers that
want the rasteriser's raw products the way ``MeshRasterizer`` returns them (render_model/mano_layer.py:1083,
metric/meshLoss.py builds ``Meshes`` the same way).

Differences from pytorch3d, all forced by what the kernels are for: the topology is the handle's shared face
list (no packing of per-mesh face lists), the sample grid is the per-hand crop grid of ``dsf_view_setup`` instead
of a fixed S x S NDC raster, K = faces_per_pixel = 1 and blur_radius = 0 (the only settings DSF uses)."""
from __future__ import annotations

from typing import NamedTuple

import torch

from . import _lib as L


class Fragments(NamedTuple):
    pix_to_face: torch.Tensor    # (N,R,R,1) int64, index into the packed face list (mesh * F + face), -1 background
    zbuf: torch.Tensor           # (N,R,R,1) view-space depth, -1 background
    bary_coords: torch.Tensor    # (N,R,R,1,3), -1 background
    dists: torch.Tensor          # (N,R,R,1) signed squared NDC distance to the nearest edge, -1 background


def rasterize_meshes(layer, verts_cam, view, xs, ys, perspective_correct=False):
    """verts_cam (N,779,3) camera-space mm; view / xs / ys from dsf_view_setup (Render._view).
    -> Fragments with pytorch3d's shapes, dtypes and packed face indices."""
    lib = L.lib()
    verts_cam = L.f32c(verts_cam.detach())
    N, R = verts_cam.shape[0], xs.shape[1]
    dev = verts_cam.device
    img = torch.empty(N, R, R, device=dev)
    p2f = torch.empty(N, R, R, dtype=torch.int32, device=dev)
    zbuf = torch.empty(N, R, R, device=dev)
    bary = torch.empty(N, R, R, 3, device=dev)
    dists = torch.empty(N, R, R, device=dev)
    L.check(lib.dsf_raster_forward(layer._handle, N, verts_cam.data_ptr(), view.data_ptr(), xs.data_ptr(), ys.data_ptr(),
                                   R, img.data_ptr(), p2f.data_ptr(), zbuf.data_ptr(), bary.data_ptr(), dists.data_ptr(),
                                   None, 0.0, None, L.RASTER_PERSPECTIVE_CORRECT if perspective_correct else 0,
                                   L.stream_ptr()))
    F = int(layer.faces_int.shape[0])
    packed = p2f.long()
    packed = torch.where(packed >= 0, packed + torch.arange(N, device=dev).view(N, 1, 1) * F, packed)
    return Fragments(packed.unsqueeze(-1), zbuf.unsqueeze(-1), bary.unsqueeze(-2), dists.unsqueeze(-1))


class MeshRasterizer:
    """``MeshRasterizer(cameras, raster_settings)(meshes)`` of the reference (:952, :1083) for a ``Render``:
    call with the placed vertices and the per-hand crop (center3d, cube) -> Fragments."""

    def __init__(self, render):
        self.render = render

    def __call__(self, verts_cam, center3d, cube_size, M=None):
        view, xs, ys, _ = self.render._view(center3d, cube_size, M)
        return rasterize_meshes(self.render.mano_layer, verts_cam, view, xs, ys, self.render.perspective_correct)
