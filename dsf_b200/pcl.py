"""Depth crop -> point cloud behind the reference's signatures (methods of the data loader there):
``Img2pcl`` (data/render_loader.py:1121-1156) and ``uvdImg2xyzImg`` (:1190-1200).  One CUDA kernel
each (dsf_img2pcl, dsf_uvd_img_to_xyz); no per-hand Python loop, no host synchronisation."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

NYU_PARAS = (588.03, 587.07, 320.0, 240.0)


def _intr(paras):
    return (C.c_float * 4)(*[float(v) for v in paras])


def Img2pcl(img, feature_size, center, M, cube, sample_num=1024, paras=NYU_PARAS, img_size=None, flip=1.0,
            seed=0, return_count=False):
    """Same arguments as ``loader.Img2pcl(img, feature_size, center, M, cube, sample_num)`` plus the
    loader's ``paras`` / ``img_size`` / ``flip`` attributes and the seed of the draw.

    Returns (B, sample_num, 3) cube-normalised points: the foreground list (pixel order) repeated
    ``sample_num // n`` times followed by ``sample_num % n`` foreground points drawn without
    replacement - the layout of :1141-1153; an empty crop gives zeros (:1144).  ``sample_num=0``
    returns the padded (B, feature_size**2, 3) list of all foreground points together with the
    per-hand count (the reference returns a ragged list there)."""
    lib = L.lib()
    img_c = L.f32c(img.detach())
    R = img_c.shape[-1]
    if img_c.shape[-2] != R:
        raise ValueError("square images only")
    B = img_c.numel() // (R * R)
    center, M, cube = L.f32c(center), L.f32c(M), L.f32c(cube)
    rows = sample_num if sample_num > 0 else feature_size * feature_size
    pcl = torch.empty(B, rows, 3, dtype=torch.float32, device=img_c.device)
    count = torch.empty(B, dtype=torch.int32, device=img_c.device)
    L.check(lib.dsf_img2pcl(B, R, int(feature_size), img_c.data_ptr(), center.data_ptr(), cube.data_ptr(),
                            M.data_ptr(), _intr(paras), float(R if img_size is None else img_size), float(flip),
                            int(sample_num), int(seed) & (2 ** 64 - 1), pcl.data_ptr(), count.data_ptr(),
                            L.stream_ptr()))
    if sample_num == 0 or return_count:
        return pcl, count
    return pcl


def uvdImg2xyzImg(uvd_img, center, M, cube, paras=NYU_PARAS, img_size=None, flip=1.0):
    """``loader.uvdImg2xyzImg(uvd_img, center, M, cube)`` -> (xyz_img mm, xyz_img_normal), (B,3,R,R)."""
    lib = L.lib()
    img_c = L.f32c(uvd_img.detach())
    R = img_c.shape[-1]
    B = img_c.numel() // (R * R)
    center, M, cube = L.f32c(center), L.f32c(M), L.f32c(cube)
    xyz = torch.empty(B, 3, R, R, dtype=torch.float32, device=img_c.device)
    xyz_n = torch.empty_like(xyz)
    L.check(lib.dsf_uvd_img_to_xyz(B, R, img_c.data_ptr(), center.data_ptr(), cube.data_ptr(), M.data_ptr(),
                                   _intr(paras), float(R if img_size is None else img_size), float(flip),
                                   xyz.data_ptr(), xyz_n.data_ptr(), L.stream_ptr()))
    return xyz, xyz_n


def target_from_u16(depth_mm, center, cube, invalid_value=0, out=None):
    """Cropped sensor depth (B,R,R) uint16 millimetres -> the normalised fp32 target, on the device
    (``loader.normalize_img``, data/render_loader.py:738-745, which the reference runs on the CPU)."""
    lib = L.lib()
    if depth_mm.dtype != torch.uint16 or not depth_mm.is_cuda or not depth_mm.is_contiguous():
        raise ValueError("depth_mm must be a contiguous CUDA uint16 tensor")
    R = depth_mm.shape[-1]
    B = depth_mm.numel() // (R * R)
    center, cube = L.f32c(center), L.f32c(cube)
    if out is None:
        out = torch.empty(B, R, R, dtype=torch.float32, device=depth_mm.device)
    L.check(lib.dsf_target_from_u16(B, R, depth_mm.data_ptr(), center.data_ptr(), cube.data_ptr(),
                                    int(invalid_value), out.data_ptr(), L.stream_ptr()))
    return out
