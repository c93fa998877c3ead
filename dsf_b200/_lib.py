"""ctypes binding of libdsf_b200.so (the C ABI declared in include/dsf_b200.h).

There is deliberately no CPU or pure-torch fallback: if the shared library is missing, or no CUDA
device is visible, every op raises.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("DSF_B200_LIB", os.path.join(_HERE, "libdsf_b200.so"))   # override: kernel-tuning experiments

VIEW_STRIDE = 20
RASTER_PERSPECTIVE_CORRECT = 1   # DSF_RASTER_PERSPECTIVE_CORRECT
RASTER_SEPARATE_BACKWARD = 2     # DSF_RASTER_SEPARATE_BACKWARD
NV, NVW, NJ, NJOUT, NSPHERE = 778, 779, 16, 21, 66

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


class DsfManoHost(C.Structure):
    _fields_ = [
        ("v_template", c_float_p), ("shapedirs", c_float_p), ("posedirs", c_float_p),
        ("j_regressor", c_float_p), ("hands_comp", c_float_p), ("hands_mean", c_float_p),
        ("weights", c_float_p), ("parents", c_int_p), ("faces", c_int_p), ("n_faces", C.c_int),
    ]


class DsfManoParams(C.Structure):
    _fields_ = [
        ("quat", C.c_void_p), ("ld_quat", C.c_int), ("quat_dim", C.c_int),
        ("theta", C.c_void_p), ("ld_theta", C.c_int), ("ncomp", C.c_int),
        ("beta", C.c_void_p), ("ld_beta", C.c_int),
        ("cam", C.c_void_p), ("ld_cam", C.c_int),
    ]


class DsfManoGrads(C.Structure):
    _fields_ = [
        ("quat", C.c_void_p), ("ld_quat", C.c_int),
        ("theta", C.c_void_p), ("ld_theta", C.c_int),
        ("beta", C.c_void_p), ("ld_beta", C.c_int),
        ("cam", C.c_void_p), ("ld_cam", C.c_int),
    ]


# name -> (restype, argtypes); every symbol include/dsf_b200.h declares
_VP, _I, _F, _L = C.c_void_p, C.c_int, C.c_float, C.c_long
SIGNATURES = {
    "dsf_last_error_string": (C.c_char_p, []),
    "dsf_version": (_I, []),
    "dsf_last_launch_count": (_I, []),
    "dsf_mano_create": (_I, [C.POINTER(DsfManoHost), C.POINTER(_VP)]),
    "dsf_mano_free": (_I, [_VP]),
    "dsf_mano_workspace_floats": (_L, [_I]),
    "dsf_mano_faces_device": (_VP, [_VP, c_int_p]),
    "dsf_mano_forward": (_I, [_VP, _I, C.POINTER(DsfManoParams), _F, _VP, _VP, _VP, _VP, _VP]),
    "dsf_mano_backward": (_I, [_VP, _I, C.POINTER(DsfManoParams), _F, _VP, _VP, _VP, _VP,
                               C.POINTER(DsfManoGrads), _VP, _VP]),
    "dsf_view_setup": (_I, [_I, _I, _VP, _VP, c_float_p, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP]),
    "dsf_raster_forward": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _F, _VP, _I, _VP]),
    "dsf_raster_tiles": (_I, [_I]),
    "dsf_raster_backward": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _I, _VP, _VP, _VP, _I, _VP]),
    "dsf_depth_loss": (_I, [_I, _I, _I, _VP, _VP, _F, _F, _VP, _VP, _VP, _VP]),
    "dsf_coll_forward_backward": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP]),
    "dsf_point_face_forward": (_I, [_I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "dsf_point_face_stats": (_I, [_I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "dsf_point_face_backward": (_I, [_I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "dsf_sphere_set": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP]),
    "dsf_seg_pcl": (_I, [_I, _I, _VP, _VP, _VP, _VP, _VP]),
    "dsf_joint_icp_forward": (_I, [_I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "dsf_joint_icp_backward": (_I, [_I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "dsf_fit_workspace_floats": (_L, [_I, _I]),
    "dsf_fit_step": (_I, [_VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _F, _I, _VP, _I, _VP, c_float_p,
                          _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "dsf_fit_step_rows": (_I, [_VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _F, _I, _VP, _I, _VP, c_float_p,
                               _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "dsf_raster_loss_workspace_floats": (_L, [_I, _I]),
    "dsf_raster_loss_grad": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _I, _VP, _F, _F, _I, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "dsf_sum_totals": (_I, [_I, _VP, _I, _VP, _VP]),
    "dsf_fit_views_workspace_floats": (C.c_long, [_I, _I, _I]),
    "dsf_fit_step_views": (_I, [_VP, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _F, _VP, _VP, _VP, _VP, _VP,
                                _VP, _VP, _VP, _I, _VP]),
    "dsf_render_workspace_floats": (C.c_long, [_I]),
    "dsf_render_forward": (_I, [_VP, _I, _I, _VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, c_float_p, _VP, _VP, _VP, _VP,
                                _VP, _VP, _VP, _VP, _I, _VP]),
    "dsf_render_backward": (_I, [_VP, _I, _I, _VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, c_float_p, _VP, _VP, _VP, _VP,
                                 _VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "dsf_crop_hand": (_I, [_I, _I, _VP, _VP, _I, _VP, _VP, _VP, c_float_p, _F, _F, _F, _VP, _VP, _VP]),
    "dsf_img2pcl": (_I, [_I, _I, _I, _VP, _VP, _VP, _VP, c_float_p, _F, _F, _I, C.c_ulonglong, _VP, _VP, _VP]),
    "dsf_target_from_u16": (_I, [_I, _I, _VP, _VP, _VP, _I, _VP, _VP]),
    "dsf_target_from_u16_rows": (_I, [_I, _I, _VP, _VP, _VP, _VP, _VP, _I, _VP, _VP]),
    "dsf_pack_u16_rows": (_L, [_I, _I, _VP, _VP, _VP, _I, _VP, _VP, _VP, _L]),
    "dsf_intersect_workspace_bytes": (C.c_long, [_I, _I, _I, _I]),
    "dsf_intersect_vox": (_I, [_I, _I, _VP, _I, _VP, _VP, _I, _VP, _VP, _VP, C.c_double, _VP, _VP, _VP, _VP, _VP, _VP]),
    "dsf_rotate_points": (_I, [_I, _I, _VP, _VP, _VP, _VP, _VP]),
    "dsf_rotate_points_backward": (_I, [_I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "dsf_mask_img": (_I, [_I, _I, _VP, _I, _VP, _VP, _VP, _VP]),
    "dsf_synth2real": (_I, [_I, _I, _VP, _VP, _I, _F, _F, _VP, _VP]),
    "dsf_chamfer_forward": (_I, [_I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "dsf_chamfer_backward": (_I, [_I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "dsf_uvd_img_to_xyz": (_I, [_I, _I, _VP, _VP, _VP, _VP, c_float_p, _F, _F, _VP, _VP, _VP]),
}

_lib = None


def load_library() -> C.CDLL:
    """dlopen the library and bind every declared symbol (no device needed for this step)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(dsf_b200 has no CPU fallback)")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def lib() -> C.CDLL:
    """Library for compute calls: requires a CUDA device."""
    if not torch.cuda.is_available():
        raise RuntimeError("dsf_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return load_library()


def check(rc: int) -> None:
    if rc != 0:
        msg = load_library().dsf_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"libdsf_b200 error {rc}: {msg}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def f32c(t: torch.Tensor) -> torch.Tensor:
    """float32, contiguous, on the current CUDA device."""
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_cuda:
        t = t.cuda()
    return t.contiguous()


def rows(t: torch.Tensor):
    """(tensor, leading dimension) for a 2-D float32 CUDA tensor with unit inner stride; column
    slices of a (B,62) parameter tensor are passed without a copy."""
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_cuda:
        t = t.cuda()
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D tensor, got shape {tuple(t.shape)}")
    if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t, t.stride(0)
