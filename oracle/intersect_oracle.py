"""TEST INFRASTRUCTURE ONLY - ctypes front end of oracle/intersect_oracle.c (row I1; PARITY UNPINNED,
see the header of the C file)."""
import ctypes as C

import numpy as np

from .raster_oracle import lib as _lib


def intersect_vox(verts, topo, pitch=2.0):
    """verts (B,n_verts,3) float32 numpy; topo: any object with the PartTopology arrays.
    Returns (volume (B,) f64, pair_counts (B,P,P) i64, voxel_counts (B,P) i64)."""
    lib = _lib()
    v = np.ascontiguousarray(verts, np.float32)
    B, P = v.shape[0], topo.n_parts
    pc = np.zeros((B, P, P), np.int64)
    vc = np.zeros((B, P), np.int64)
    vol = np.zeros(B, np.float64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    cap_ptr, cap_idx = np.ascontiguousarray(topo.cap_ptr, np.int32), np.ascontiguousarray(topo.cap_idx, np.int32)
    part_ptr, faces = np.ascontiguousarray(topo.part_ptr, np.int32), np.ascontiguousarray(topo.faces, np.int32)
    mask = np.ascontiguousarray(topo.pair_mask, np.uint8)
    fn = lib.orc_batch_intersect_vox
    fn.restype = None
    fn.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                   C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    fn(B, p(v), v.shape[1], len(topo.cap_loops), p(cap_ptr), p(cap_idx), P, p(part_ptr), p(faces), p(mask),
       float(pitch), p(pc), p(vc), p(vol))
    return vol, pc, vc
