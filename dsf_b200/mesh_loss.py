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
        order = torch.empty(B * (P + F), dtype=torch.int32, device=dev)  # scratch: spatial order of points and faces
        L.check(lib.dsf_point_face_forward(B, P, V, F, points.data_ptr(), verts.data_ptr(), faces_i32.data_ptr(),
                                           dists.data_ptr(), idxs.data_ptr(), order.data_ptr(), L.stream_ptr()))
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


class _SubsetPointFaceDistance(torch.autograd.Function):
    """Every point against the face subset of its own label only (dsf_joint_icp_forward/backward)."""

    @staticmethod
    def forward(ctx, points, verts, seg_i32, sub_ptr, sub_faces):
        lib = L.lib()
        points = L.f32c(points)
        verts = L.f32c(verts)
        B, P, _ = points.shape
        dev = points.device
        dists = torch.empty(B, P, device=dev)
        idxs = torch.empty(B, P, dtype=torch.int32, device=dev)
        L.check(lib.dsf_joint_icp_forward(B, P, verts.shape[1], sub_ptr.numel() - 1, points.data_ptr(),
                                          verts.data_ptr(), seg_i32.data_ptr(), sub_ptr.data_ptr(),
                                          sub_faces.data_ptr(), dists.data_ptr(), idxs.data_ptr(), L.stream_ptr()))
        ctx.save_for_backward(points, verts, seg_i32, sub_ptr, sub_faces, idxs)
        return dists

    @staticmethod
    def backward(ctx, g):
        lib = L.lib()
        points, verts, seg_i32, sub_ptr, sub_faces, idxs = ctx.saved_tensors
        B, P, _ = points.shape
        gp = torch.empty_like(points)
        gv = torch.empty_like(verts)
        g = L.f32c(g)
        L.check(lib.dsf_joint_icp_backward(B, P, verts.shape[1], points.data_ptr(), verts.data_ptr(),
                                           seg_i32.data_ptr(), sub_ptr.data_ptr(), sub_faces.data_ptr(),
                                           idxs.data_ptr(), g.data_ptr(), gp.data_ptr(), gv.data_ptr(),
                                           L.stream_ptr()))
        return gp, gv, None, None, None


_SUBSET_CACHE = {}


def _subset_csr(faces_list, device):
    key = (tuple(int(f.data_ptr()) for f in faces_list), tuple(int(f.shape[0]) for f in faces_list), str(device))
    hit = _SUBSET_CACHE.get(key)
    if hit is None:
        sizes = [int(f.shape[0]) for f in faces_list]
        ptr = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int32, device=device)
        faces = torch.cat([_faces_i32(f, device) for f in faces_list], 0).contiguous()
        hit = (ptr, faces)
        _SUBSET_CACHE[key] = hit
    return hit


def _part_loss(mesh, pcl, faces_list, pcl_seg):
    """meshLoss.py:387-394 for all parts at once: the reference replicates mesh and cloud once per
    part and masks afterwards; here each point only ever meets the faces of its own part."""
    n = len(faces_list)
    ptr, faces = _subset_csr(faces_list, pcl.device)
    seg = pcl_seg.to(torch.int32)
    seg = torch.where((seg >= 1) & (seg <= n), seg, torch.zeros_like(seg)).contiguous()
    d = _SubsetPointFaceDistance.apply(pcl, mesh, seg, ptr, faces)                     # (B,P)
    onehot = torch.nn.functional.one_hot(seg.long(), n + 1)[..., 1:].to(d.dtype)        # (B,P,n)
    total = torch.einsum("bp,bpk->bk", d, onehot)
    valid = torch.einsum("bp,bpk->bk", d.gt(0).to(d.dtype), onehot)
    loss = total / (valid + 1e-8)
    return torch.where(valid.eq(0), torch.zeros_like(loss), loss)


def JointICPLoss(mesh, pcl, faces, pcl_seg):
    """15 per-joint face subsets (MANO_SMPL.joint_faces); points gated by pcl_seg == k+1."""
    return _part_loss(mesh, pcl, faces, pcl_seg)


def FingerICPLoss(mesh, pcl, faces, pcl_seg):
    """5 per-finger face subsets (MANO_SMPL.finger_faces)."""
    return _part_loss(mesh, pcl, faces, pcl_seg)
