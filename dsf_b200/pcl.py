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


class RowRunTarget:
    """Host-side row-run packing of sensor crops (the loader's output format for dsf_target_from_u16_rows): per row
    only the span from the first to the last non-background pixel.  rows (B,R,2) uint16 = (first column, length),
    hand_offset (B+1) int32 (pixels), payload (n,) uint16 - pinned CPU tensors unless ``pin=False``."""

    def __init__(self, rows, hand_offset, payload, R):
        self.rows, self.hand_offset, self.payload, self.R = rows, hand_offset, payload, int(R)

    @property
    def batch(self):
        return self.rows.shape[0]

    @property
    def nbytes(self):
        return self.rows.numel() * 2 + self.hand_offset.numel() * 4 + self.payload.numel() * 2


def pack_target_rows(depth_mm, center, cube, invalid_value=0, pin=True):
    """(B,R,R) uint16 millimetre crops (CPU tensor or ndarray) -> RowRunTarget.  Runs the library's host packer
    (dsf_pack_u16_rows; no device involved): this is work for the data loader, once per sample."""
    import numpy as np

    lib = L.load_library()
    d = depth_mm.cpu().numpy() if isinstance(depth_mm, torch.Tensor) else np.asarray(depth_mm)
    if d.dtype != np.uint16 or d.ndim != 3 or d.shape[1] != d.shape[2]:
        raise ValueError("depth_mm must be (B,R,R) uint16")
    d = np.ascontiguousarray(d)
    B, R = d.shape[0], d.shape[1]
    c = np.ascontiguousarray(center.cpu().numpy() if isinstance(center, torch.Tensor) else center, dtype=np.float32)
    q = np.ascontiguousarray(cube.cpu().numpy() if isinstance(cube, torch.Tensor) else cube, dtype=np.float32)
    rows = np.empty((B, R, 2), np.uint16)
    off = np.empty(B + 1, np.uint32)
    payload = np.empty(B * R * R, np.uint16)
    n = lib.dsf_pack_u16_rows(B, R, d.ctypes.data, c.ctypes.data, q.ctypes.data, int(invalid_value), rows.ctypes.data,
                              off.ctypes.data, payload.ctypes.data, payload.size)
    if n < 0:
        raise RuntimeError("dsf_pack_u16_rows failed (bad arguments)")
    t_rows = torch.from_numpy(rows)
    t_off = torch.from_numpy(off.view(np.int32))
    t_pay = torch.from_numpy(payload[:max(int(n), 1)].copy())
    if pin and torch.cuda.is_available():
        t_rows, t_off, t_pay = t_rows.pin_memory(), t_off.pin_memory(), t_pay.pin_memory()
    return RowRunTarget(t_rows, t_off, t_pay, R)


def target_from_u16_rows(packed, center, cube, invalid_value=0, out=None, buffers=None):
    """RowRunTarget (host) -> normalised fp32 target (B,R,R) on the device: three small H2D copies + one kernel.
    ``buffers`` = optional persistent device tensors (rows, hand_offset, payload) to copy into."""
    lib = L.lib()
    center, cube = L.f32c(center), L.f32c(cube)
    dev = center.device
    B, R = packed.batch, packed.R
    if buffers is None:
        buffers = (torch.empty(B, R, 2, dtype=torch.uint16, device=dev), torch.empty(B + 1, dtype=torch.int32, device=dev),
                   torch.empty(B * R * R, dtype=torch.uint16, device=dev))
    d_rows, d_off, d_pay = buffers
    n = packed.payload.numel()
    d_rows.copy_(packed.rows, non_blocking=True)
    d_off.copy_(packed.hand_offset, non_blocking=True)
    d_pay[:n].copy_(packed.payload, non_blocking=True)
    if out is None:
        out = torch.empty(B, R, R, dtype=torch.float32, device=dev)
    L.check(lib.dsf_target_from_u16_rows(B, R, d_rows.data_ptr(), d_off.data_ptr(), d_pay.data_ptr(), center.data_ptr(),
                                         cube.data_ptr(), int(invalid_value), out.data_ptr(), L.stream_ptr()))
    return out
