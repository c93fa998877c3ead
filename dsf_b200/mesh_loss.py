"""Point-cloud to mesh distance losses behind the reference's signatures (metric/meshLoss.py).

``ICPLoss(mesh, pcl, faces)`` (:347-353), ``JointICPLoss(mesh, pcl, faces_list, pcl_seg)`` (:377-394)
and ``FingerICPLoss`` (:356-374) call one CUDA kernel per face list (dsf_point_face_forward /
_backward) instead of packing pytorch3d Meshes / Pointclouds and replicating the mesh 15 times.
"""
from __future__ import annotations

import torch

from . import _lib as L


class _PointFaceDistance(torch.autograd.Function):
    """meshLoss.py:21-70 with the batch-shared face list DSF always uses."""

    @staticmethod
    def forward(ctx, points, verts, faces_i32):
        lib = L.lib()
        points = L.f32c(points)
        verts = L.f32c(verts)
        B, P, _ = points.shape
        V = verts.shape[1]
        F = faces_i32.shape[0]
        dev = points.device
        dists = torch.empty(B, P, device=dev)
        idxs = torch.empty(B, P, dtype=torch.int32, device=dev)
        L.check(lib.dsf_point_face_forward(B, P, V, F, points.data_ptr(), verts.data_ptr(), faces_i32.data_ptr(),
                                           dists.data_ptr(), idxs.data_ptr(), L.stream_ptr()))
        ctx.save_for_backward(points, verts, faces_i32, idxs)
        ctx.mark_non_differentiable(idxs)
        return dists, idxs

    @staticmethod
    def backward(ctx, g_dists, _g_idx):
        lib = L.lib()
        points, verts, faces_i32, idxs = ctx.saved_tensors
        B, P, _ = points.shape
        V = verts.shape[1]
        g_dists = L.f32c(g_dists)
        gp = torch.empty_like(points)
        gv = torch.empty_like(verts)
        L.check(lib.dsf_point_face_backward(B, P, V, faces_i32.shape[0], points.data_ptr(), verts.data_ptr(),
                                            faces_i32.data_ptr(), idxs.data_ptr(), g_dists.data_ptr(),
                                            gp.data_ptr(), gv.data_ptr(), L.stream_ptr()))
        return gp, gv, None


def _faces_i32(faces: torch.Tensor, device) -> torch.Tensor:
    if faces.dim() != 2 or faces.shape[1] != 3:
        raise ValueError("faces must be (F,3)")          # cf. the ValueError at meshLoss.py:249-250
    return faces.to(device=device, dtype=torch.int32).contiguous()


def point_face_distance(points, verts, faces):
    """(B,P,3), (B,V,3), (F,3) -> squared distance of every point to its closest face, (B,P)."""
    return _PointFaceDistance.apply(points, verts, _faces_i32(faces, points.device))[0]


def ICPLoss(mesh, pcl, faces):
    if mesh.shape[0] != pcl.shape[0]:
        raise ValueError("meshes and pointclouds must be equal sized batches")
    return point_face_distance(pcl, mesh, faces).mean(-1)


def _part_loss(mesh, pcl, faces_list, pcl_seg):
    out = []
    for k, faces in enumerate(faces_list):
        d = point_face_distance(pcl, mesh, faces)
        d = torch.where(pcl_seg.eq(k + 1), d, torch.zeros_like(d))
        valid = d.gt(0).sum(-1)
        loss = d.sum(-1) / (valid + 1e-8)
        out.append(torch.where(valid.eq(0), torch.zeros_like(loss), loss))
    return torch.stack(out, dim=-1)


def JointICPLoss(mesh, pcl, faces, pcl_seg):
    """15 per-joint face subsets (MANO_SMPL.joint_faces); points gated by pcl_seg == k+1."""
    return _part_loss(mesh, pcl, faces, pcl_seg)


def FingerICPLoss(mesh, pcl, faces, pcl_seg):
    """5 per-finger face subsets (MANO_SMPL.finger_faces)."""
    return _part_loss(mesh, pcl, faces, pcl_seg)
