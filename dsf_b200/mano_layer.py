"""Drop-in host side of DSF's hand layer and differentiable depth renderer.

Mirrors the call signatures of the reference's ``render_model/mano_layer.py``:

* ``MANO_SMPL(mano_pkl_path, dataset, scale=1000)`` with ``forward`` (:573), ``get_mano_vertices``
  (:643) and ``calculate_coll`` (:373).  ``ManoLayer`` is an alias (the name BASELINE.json uses).
* ``Render(mano_path, dataset, cam_para, image_size, crop_size)`` with ``render`` (:1071),
  ``forward`` (:983), ``M_render`` (:1100), ``normal_render`` (:1042), ``getDepth`` (:1204),
  ``mesh2img`` (:1190), ``get_mesh_xyz`` (:1171) and the small camera helpers.

All heavy work (MANO forward/backward, rasterisation forward/backward, collision) runs in
hand-written sm_100a kernels behind the C ABI of ``include/dsf_b200.h``; torch supplies device
memory, streams and the autograd graph edges.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import pickle
import random

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .mano_extras import SphereVariants

# joint re-orderings from MANO to each dataset's skeleton (same tables as the reference, :36-81)
MANO2HANDS = [0, 13, 1, 4, 10, 7, 14, 15, 20, 2, 3, 16, 5, 6, 17, 11, 12, 19, 8, 9, 18]
MANO2MSRA = [0, 1, 2, 3, 16, 4, 5, 6, 17, 10, 11, 12, 19, 7, 8, 9, 18, 13, 14, 15, 20]
MANO2ICVL = [0, 13, 14, 15, 1, 2, 3, 4, 5, 6, 10, 11, 12, 7, 8, 9]
MANO2NYU = [18, 8, 19, 11, 17, 5, 16, 2, 20, 15, 14, 0]
HANDS2MANO = [0, 2, 9, 10, 3, 12, 13, 5, 18, 19, 4, 15, 16, 1, 6, 7, 11, 14, 20, 17, 8]

WRIST_RING = [121, 214, 215, 279, 239, 234, 92, 38, 122, 118, 117, 119, 120, 108, 79, 78]
TIP_VERTS = [333, 444, 672, 555, 744]


def _np32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).astype(np.float32))


def _fptr(a):
    return a.ctypes.data_as(L.c_float_p)


# ------------------------------------------------------------------------------------------------
# autograd edges
# ------------------------------------------------------------------------------------------------
class _ManoFunction(torch.autograd.Function):
    """MANO forward/backward through dsf_mano_forward / dsf_mano_backward."""

    @staticmethod
    def forward(ctx, layer, quat, theta, beta, cam, unit_scale):
        lib = L.lib()
        quat, ldq = L.rows(quat)
        theta, ldt = L.rows(theta)
        beta, ldb = L.rows(beta)
        B = beta.shape[0]
        if cam is not None:
            cam, ldc = L.rows(cam)
        else:
            ldc = 0
        p = L.DsfManoParams(quat.data_ptr(), ldq, quat.shape[1], theta.data_ptr(), ldt, theta.shape[1],
                            beta.data_ptr(), ldb, L.ptr(cam), ldc)
        dev = beta.device
        verts = torch.empty(B, L.NVW, 3, device=dev)
        joints = torch.empty(B, L.NJOUT, 3, device=dev)
        Rs = torch.empty(B, 15, 3, 3, device=dev)
        ws = torch.empty(lib.dsf_mano_workspace_floats(B), device=dev)
        L.check(lib.dsf_mano_forward(layer._handle, B, C.byref(p), float(unit_scale), verts.data_ptr(),
                                     joints.data_ptr(), Rs.data_ptr(), ws.data_ptr(), L.stream_ptr()))
        ctx.layer = layer
        ctx.unit_scale = float(unit_scale)
        ctx.has_cam = cam is not None
        ctx.save_for_backward(quat, theta, beta, cam if cam is not None else torch.empty(0, device=dev),
                              verts, joints, ws)
        ctx.mark_non_differentiable(Rs)
        return verts, joints, Rs

    @staticmethod
    def backward(ctx, g_verts, g_joints, _g_rs):
        lib = L.lib()
        quat, theta, beta, cam, verts, joints, ws = ctx.saved_tensors
        cam = cam if ctx.has_cam else None
        B = beta.shape[0]
        quat, ldq = L.rows(quat)
        theta, ldt = L.rows(theta)
        beta, ldb = L.rows(beta)
        ldc = 0
        if cam is not None:
            cam, ldc = L.rows(cam)
        p = L.DsfManoParams(quat.data_ptr(), ldq, quat.shape[1], theta.data_ptr(), ldt, theta.shape[1],
                            beta.data_ptr(), ldb, L.ptr(cam), ldc)
        dev = beta.device
        gq = torch.empty(B, quat.shape[1], device=dev)
        gt = torch.empty(B, theta.shape[1], device=dev)
        gb = torch.empty(B, 10, device=dev)
        gc = torch.empty(B, 4, device=dev) if cam is not None else None
        g = L.DsfManoGrads(gq.data_ptr(), gq.shape[1], gt.data_ptr(), gt.shape[1], gb.data_ptr(), 10,
                           L.ptr(gc), 4)
        gv = L.f32c(g_verts) if g_verts is not None else None
        gj = L.f32c(g_joints) if g_joints is not None else None
        L.check(lib.dsf_mano_backward(ctx.layer._handle, B, C.byref(p), ctx.unit_scale, verts.data_ptr(),
                                      joints.data_ptr(), L.ptr(gv), L.ptr(gj), C.byref(g), ws.data_ptr(),
                                      L.stream_ptr()))
        return None, gq, gt, gb, gc, None


class _RasterFunction(torch.autograd.Function):
    """Normalised depth image from camera-space vertices (dsf_raster_forward / _backward)."""

    @staticmethod
    def forward(ctx, layer, verts_cam, view, xs, ys, flags=0):
        lib = L.lib()
        verts_cam = L.f32c(verts_cam)
        NM = verts_cam.shape[0]
        R = xs.shape[1]
        dev = verts_cam.device
        img = torch.empty(NM, 1, R, R, device=dev)
        p2f = torch.empty(NM, R, R, dtype=torch.int32, device=dev)
        L.check(lib.dsf_raster_forward(layer._handle, NM, verts_cam.data_ptr(), view.data_ptr(), xs.data_ptr(),
                                       ys.data_ptr(), R, img.data_ptr(), p2f.data_ptr(), None, None, None,
                                       None, 0.0, None, flags, L.stream_ptr()))
        ctx.layer = layer
        ctx.flags = flags
        ctx.save_for_backward(verts_cam, view, xs, ys, p2f)
        ctx.mark_non_differentiable(p2f)
        return img, p2f

    @staticmethod
    def backward(ctx, g_img, _g_p2f):
        lib = L.lib()
        verts_cam, view, xs, ys, p2f = ctx.saved_tensors
        NM = verts_cam.shape[0]
        R = xs.shape[1]
        g_img = L.f32c(g_img)
        gv = torch.empty_like(verts_cam)
        L.check(lib.dsf_raster_backward(ctx.layer._handle, NM, verts_cam.data_ptr(), view.data_ptr(),
                                        xs.data_ptr(), ys.data_ptr(), R, p2f.data_ptr(), g_img.data_ptr(),
                                        gv.data_ptr(), ctx.flags, L.stream_ptr()))
        return None, gv, None, None, None, None


class _CollFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layer, joints, mesh):
        lib = L.lib()
        joints = L.f32c(joints)
        mesh = L.f32c(mesh)
        B = joints.shape[0]
        dev = joints.device
        out = torch.empty(1, device=dev)
        per_hand = torch.empty(B, 2, device=dev)
        gj = torch.empty_like(joints)
        L.check(lib.dsf_coll_forward_backward(layer._handle, B, joints.data_ptr(), mesh.data_ptr(),
                                              out.data_ptr(), per_hand.data_ptr(), gj.data_ptr(),
                                              L.stream_ptr()))
        ctx.save_for_backward(gj)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (gj,) = ctx.saved_tensors
        return None, gj * g, None


# ------------------------------------------------------------------------------------------------
# MANO layer
# ------------------------------------------------------------------------------------------------
class MANO_SMPL(SphereVariants, nn.Module):
    """B200-native replacement of the reference ``MANO_SMPL`` (mano_layer.py:82-770)."""

    def __init__(self, mano_pkl_path, dataset, scale=1000):
        super().__init__()
        if "msra" in dataset:
            self.transfer = MANO2MSRA
        elif "icvl" in dataset:
            self.transfer = MANO2ICVL
        elif "hands" in dataset:
            self.transfer = MANO2HANDS
        elif "nyu" in dataset:
            self.transfer = MANO2NYU
        else:
            self.transfer = range(21)
        self.dataset = dataset
        self.scale = scale
        if isinstance(mano_pkl_path, dict):
            model = mano_pkl_path
        else:
            with open(mano_pkl_path, "rb") as f:
                model = pickle.load(f, encoding="latin1")
        self._init_constants(model)

    # -- constants in the reference layouts (M0) ---------------------------------------------------
    def _init_constants(self, model):
        if not torch.cuda.is_available():
            raise RuntimeError("dsf_b200.MANO_SMPL needs a CUDA device; there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        self.is_cuda = True
        f = np.asarray(model["f"]).astype(np.int64)
        fan = np.array([[WRIST_RING[i], WRIST_RING[(i + 1) % 16], 778] for i in range(16)])
        faces = np.concatenate([f, fan], 0)
        v_template = _np32(model["v_template"])
        self.size = [v_template.shape[0], 3]
        sd = np.asarray(model["shapedirs"], dtype=np.float64)
        self.num_betas = sd.shape[-1]
        shapedirs = _np32(sd.reshape(-1, self.num_betas).T)
        pd = np.asarray(model["posedirs"], dtype=np.float64)
        posedirs = _np32(pd.reshape(-1, pd.shape[-1]).T)
        jr = model["J_regressor"]
        jr = jr.toarray() if hasattr(jr, "toarray") else np.asarray(jr)
        jr16 = _np32(jr.T)                                   # (778,16)
        tips = np.zeros((778, 5), np.float32)
        tips[TIP_VERTS, np.arange(5)] = 1.0
        comp = _np32(model["hands_components"])
        mean = _np32(model["hands_mean"])
        self.parents = np.array(model["kintree_table"])[0].astype(np.int32)
        w = np.asarray(model["weights"], dtype=np.float64)
        weights = _np32(w)
        if v_template.shape != (778, 3) or shapedirs.shape != (10, 2334) or posedirs.shape != (135, 2334) \
                or jr16.shape != (778, 16) or weights.shape != (778, 16) or comp.shape != (45, 45):
            raise ValueError("MANO model arrays do not have the MANO shapes (778 vertices, 16 joints)")

        faces_i32 = np.ascontiguousarray(faces.astype(np.int32))
        parents = np.ascontiguousarray(np.where(self.parents < 0, 0, self.parents).astype(np.int32))
        host = L.DsfManoHost(_fptr(v_template), _fptr(shapedirs), _fptr(posedirs), _fptr(jr16), _fptr(comp),
                             _fptr(mean), _fptr(weights), parents.ctypes.data_as(L.c_int_p),
                             faces_i32.ctypes.data_as(L.c_int_p), faces_i32.shape[0])
        handle = C.c_void_p()
        L.check(L.lib().dsf_mano_create(C.byref(host), C.byref(handle)))
        self._handle = handle
        self._free = L.lib().dsf_mano_free

        # tensors the reference exposes as attributes / buffers
        self.faces = torch.from_numpy(faces.astype(np.float32)).to(dev)          # float, like :102-106
        self.faces_int = torch.from_numpy(faces_i32).to(dev)
        vertex_seg = np.argmax(w, axis=-1)
        self.vertex_seg = torch.from_numpy(vertex_seg.astype(np.float32))
        self.vertex_joint_index_list = [torch.from_numpy(np.nonzero(vertex_seg == k)[0]) for k in range(16)]
        vertex_joint = [np.nonzero(weights[:, k] > 0.1)[0] for k in range(16)]
        self.joint_faces = [self._faces_touching(faces, vertex_joint[k], dev) for k in range(1, 16)]
        self.vertex_finger_index_list = [
            torch.from_numpy(np.concatenate([vertex_joint[3 * k + 1], vertex_joint[3 * k + 2],
                                             vertex_joint[3 * k + 3]])) for k in range(5)]
        joint2finger = np.array([0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5])
        self.finger_seg = torch.from_numpy(joint2finger[vertex_seg])
        self.finger_faces = [self._faces_touching(faces, self.vertex_finger_index_list[k].numpy(), dev)
                             for k in range(5)]
        self.register_buffer("v_template", torch.from_numpy(v_template).to(dev))
        self.register_buffer("shapedirs", torch.from_numpy(shapedirs).to(dev))
        self.register_buffer("J_regressor", torch.from_numpy(np.concatenate([jr16, tips], 1)).to(dev))
        self.register_buffer("hands_comp", torch.from_numpy(comp).to(dev))
        self.register_buffer("hands_mean", torch.from_numpy(mean).to(dev))
        self.register_buffer("posedirs", torch.from_numpy(posedirs).to(dev))
        self.register_buffer("e3", torch.eye(3, device=dev))
        self.register_buffer("weight", torch.from_numpy(weights).to(dev))
        self.rotate_base = False
        self.child = [2, 3, 16, 5, 6, 17, 8, 9, 18, 11, 12, 19, 14, 15, 20]
        self.per_adj_shpere = 2
        self.interval = 3
        self.plam_per_adj_shpere = 4
        self.plam_interval = 5
        self.mask = self._collision_mask()

    @staticmethod
    def _faces_touching(faces, vert_ids, dev):
        sel = np.isin(faces, vert_ids).any(1)
        return torch.from_numpy(faces[sel].astype(np.float32)).to(dev)

    @staticmethod
    def _collision_mask():
        """66x66 pair table (mano_layer.py:239-269); same rules as csrc/coll.cu."""
        NP, I = 21, 3
        m = torch.zeros(66, 66)
        m[:NP, NP:] = 1.0
        m[NP:, :] = 1.0
        for b in range(15):
            root = b // 3 + 1
            rows = slice(NP + I * b, NP + I * b + I)
            if b % 3 == 0:
                m[rows, root * 4] = 0.0
                m[root * 4, rows] = 0.0
                m[rows, NP + I * b: NP + I * b + I + 3] = 0.0
            else:
                m[rows, NP + I * b - I: min(NP + I * b + 2 * I + 1, NP + 3 * I * root)] = 0.0
        th = 12 * I
        m[NP + th: NP + th + I + 1, :NP] = 0.0
        m[:NP, NP + th: NP + th + I + 1] = 0.0
        return m

    def __del__(self):
        try:
            h = self.__dict__.get("_handle")
            if h is not None and h.value:
                self.__dict__["_handle"] = None
                self.__dict__["_free"](h)
        except Exception:
            pass

    # -- M1 ------------------------------------------------------------------------------------------
    def _run(self, beta, theta, quat, cam, unit_scale):
        as_t = lambda x: x if isinstance(x, torch.Tensor) else torch.tensor(x, dtype=torch.float)
        beta, theta, quat = as_t(beta), as_t(theta), as_t(quat)
        if cam is not None:
            cam = as_t(cam)
        return _ManoFunction.apply(self, quat, theta, beta, cam, unit_scale)

    def forward(self, beta, theta, quat_or_euler, get_skin=False):
        verts, joints, Rs = self._run(beta, theta, quat_or_euler, None, 1.0)
        if get_skin:
            return verts, joints, Rs
        return joints

    # -- M4 ------------------------------------------------------------------------------------------
    def get_mano_vertices(self, quat_or_euler, pose, shape, cam, global_scale=None):
        unit = 1000.0 if global_scale is None else 1000.0 * float(global_scale)
        verts, joints, _ = self._run(shape, pose, quat_or_euler, cam, unit)
        return verts, joints

    # -- C1 ------------------------------------------------------------------------------------------
    def calculate_coll(self, joints, meshs):
        return _CollFunction.apply(self, joints, meshs.detach())

    # -- "next" row f2: sphere set and point-cloud segmentation (no gradient, as used by the trainer) --
    def _sphere_set(self, joints_c, joints_r, mesh):
        lib = L.lib()
        joints_c, joints_r, mesh = L.f32c(joints_c.detach()), L.f32c(joints_r.detach()), L.f32c(mesh.detach())
        B = joints_c.shape[0]
        c = torch.empty(B, L.NSPHERE, 3, device=joints_c.device)
        r = torch.empty(B, L.NSPHERE, device=joints_c.device)
        L.check(lib.dsf_sphere_set(self._handle, B, joints_c.data_ptr(), joints_r.data_ptr(), mesh.data_ptr(),
                                   c.data_ptr(), r.data_ptr(), L.stream_ptr()))
        return c, r

    def get_sphere_radius(self, joints, mesh):
        """(B,21,3), (B,779,3) -> 66 sphere centres (B,66,3) and radii (B,66) (mano_layer.py:271-317).
        Forward only; the differentiable use of the spheres is calculate_coll."""
        return self._sphere_set(joints, joints, mesh)

    def seg_pcl(self, joints, joints_mano, mesh, pcl):
        """mano_layer.py:404-426: label every point 0 (palm) or 1..15 (finger bone) by the nearest
        sphere surface; centres from ``joints``, radii from ``joints_mano`` and ``mesh``."""
        lib = L.lib()
        c, r = self._sphere_set(joints, joints_mano, mesh)
        pcl = L.f32c(pcl.detach())
        B, P, _ = pcl.shape
        seg = torch.empty(B, P, dtype=torch.int32, device=pcl.device)
        L.check(lib.dsf_seg_pcl(B, P, pcl.data_ptr(), c.data_ptr(), r.data_ptr(), seg.data_ptr(), L.stream_ptr()))
        return seg.long()


ManoLayer = MANO_SMPL


# ------------------------------------------------------------------------------------------------
# rigid helpers (mano_layer.py:773-893), torch glue for the rarely used view augmentation
# ------------------------------------------------------------------------------------------------
def quat2mat(quat):
    q = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([
        w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
        2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x,
        2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z], dim=1).view(-1, 3, 3)


def batch_rodrigues(theta):
    angle = torch.norm(theta + 1e-8, p=2, dim=1, keepdim=True)
    axis = theta / angle
    half = angle * 0.5
    return quat2mat(torch.cat([torch.cos(half), torch.sin(half) * axis], dim=1))


class _RenderFunction(torch.autograd.Function):
    """Render.render in one forward and one backward call (dsf_render_forward / dsf_render_backward):
    MANO, placement, rasterisation + normalisation and the three auxiliary outputs, then their adjoints
    back to the (B, 62 | 63) parameter block."""

    @staticmethod
    def forward(ctx, render, model_paras, center3d, cube_size, view, xs, ys, M):
        lib = L.lib()
        layer = render.mano_layer
        prm = L.f32c(model_paras)
        center3d, cube_size = L.f32c(center3d), L.f32c(cube_size)
        B, ld = prm.shape
        qd = 4 if ld == 63 else 3
        R = xs.shape[1]
        dev = prm.device
        img = torch.empty(B, 1, R, R, device=dev)
        p2f = torch.empty(B, R, R, dtype=torch.int32, device=dev)
        verts = torch.empty(B, L.NVW, 3, device=dev)
        joints = torch.empty(B, L.NJOUT, 3, device=dev)
        juvd, jxyz, mxyz = torch.empty_like(joints), torch.empty_like(joints), torch.empty_like(verts)
        ws = torch.empty(lib.dsf_render_workspace_floats(B), device=dev)
        L.check(lib.dsf_render_forward(layer._handle, B, R, prm.data_ptr(), ld, qd, center3d.data_ptr(),
                                       cube_size.data_ptr(), view.data_ptr(), xs.data_ptr(), ys.data_ptr(), M.data_ptr(),
                                       render._intr, img.data_ptr(), p2f.data_ptr(), verts.data_ptr(), joints.data_ptr(),
                                       juvd.data_ptr(), jxyz.data_ptr(), mxyz.data_ptr(), ws.data_ptr(),
                                       render.raster_flags, L.stream_ptr()))
        ctx.render = render
        ctx.set_materialize_grads(False)          # unused outputs arrive as None, not as zero tensors
        ctx.save_for_backward(prm, center3d, cube_size, view, xs, ys, M, verts, joints, p2f, ws)
        ctx.mark_non_differentiable(p2f)
        return img, juvd, jxyz, mxyz, p2f

    @staticmethod
    def backward(ctx, g_img, g_juvd, g_jxyz, g_mxyz, _g_p2f):
        lib = L.lib()
        prm, center3d, cube_size, view, xs, ys, M, verts, joints, p2f, ws = ctx.saved_tensors
        B, ld = prm.shape
        qd = 4 if ld == 63 else 3
        R = xs.shape[1]
        gs = [None if g is None else L.f32c(g) for g in (g_img, g_juvd, g_jxyz, g_mxyz)]
        g_prm = torch.empty_like(prm)
        L.check(lib.dsf_render_backward(ctx.render.mano_layer._handle, B, R, prm.data_ptr(), ld, qd, center3d.data_ptr(),
                                        cube_size.data_ptr(), view.data_ptr(), xs.data_ptr(), ys.data_ptr(), M.data_ptr(),
                                        ctx.render._intr, verts.data_ptr(), joints.data_ptr(), p2f.data_ptr(),
                                        L.ptr(gs[0]), L.ptr(gs[1]), L.ptr(gs[2]), L.ptr(gs[3]), g_prm.data_ptr(),
                                        ws.data_ptr(), ctx.render.raster_flags, L.stream_ptr()))
        return None, g_prm, None, None, None, None, None, None


class _RotatePoints(torch.autograd.Function):
    """out = R (p - c) + c in one kernel (dsf_rotate_points); the cotangents of the points, of R and of
    c come from one more (dsf_rotate_points_backward) - instead of B*N 3x3 GEMVs in torch.matmul."""

    @staticmethod
    def forward(ctx, pts, Rm, center):
        lib = L.lib()
        pts_c, Rm_c = L.f32c(pts), L.f32c(Rm)
        cen_c = None if center is None else L.f32c(center)
        B, n = pts_c.shape[0], pts_c.shape[1]
        out = torch.empty_like(pts_c)
        L.check(lib.dsf_rotate_points(B, n, pts_c.data_ptr(), Rm_c.data_ptr(),
                                      None if cen_c is None else cen_c.data_ptr(), out.data_ptr(), L.stream_ptr()))
        ctx.save_for_backward(pts_c, Rm_c, cen_c)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.lib()
        pts_c, Rm_c, cen_c = ctx.saved_tensors
        g = L.f32c(g)
        B, n = pts_c.shape[0], pts_c.shape[1]
        need_p, need_R, need_c = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2] and cen_c is not None
        g_p = torch.empty_like(pts_c) if need_p else None
        g_R = torch.empty_like(Rm_c) if need_R else None
        g_c = torch.empty_like(cen_c) if need_c else None
        if need_p or need_R or need_c:
            ptr = lambda t: None if t is None else t.data_ptr()
            L.check(lib.dsf_rotate_points_backward(B, n, pts_c.data_ptr(), Rm_c.data_ptr(), ptr(cen_c), g.data_ptr(),
                                                   ptr(g_p), ptr(g_R), ptr(g_c), L.stream_ptr()))
        return g_p, g_R, g_c


def RotationPoints(verts, joints, center3d, rot):
    """mano_layer.py:874-885: rotate verts (B,N,3) and joints (B,J,3) about center3d (B,3) by rot
    (axis-angle (B,3) or quaternion (B,4))."""
    rot_mat = batch_rodrigues(rot) if rot.size(-1) == 3 else quat2mat(rot)
    return _RotatePoints.apply(verts, rot_mat, center3d), _RotatePoints.apply(joints, rot_mat, center3d)


def RotationNormalPoints(points, rot):
    """mano_layer.py:888-895: pure rotation of (B,N,3) points."""
    rot_mat = batch_rodrigues(rot) if rot.size(-1) == 3 else quat2mat(rot)
    return _RotatePoints.apply(points, rot_mat, None)


# ------------------------------------------------------------------------------------------------
# renderer
# ------------------------------------------------------------------------------------------------
class Render(nn.Module):
    """B200-native replacement of the reference ``Render`` (mano_layer.py:925-1355).

    ``mode='literal'`` reproduces the reference's 640x640 raster -> resize -> crop chain by
    rasterising only the raster pixel each crop pixel reads; ``mode='direct'`` rasterises the
    crop directly with crop-space intrinsics (the benchmark configuration).

    ``perspective_correct`` is pytorch3d's RasterizationSettings flag.  The reference never passes it
    (mano_layer.py:946-950) and pytorch3d 0.4.0 - the release it pins - defaults it to False (depth
    interpolated with the screen-space barycentrics), so that is the default here; True reproduces what
    later pytorch3d releases infer for perspective cameras."""

    def __init__(self, mano_path, dataset, cam_para, image_size, crop_size=(128, 128), mode="literal",
                 perspective_correct=False):
        super().__init__()
        self.perspective_correct = bool(perspective_correct)
        self.raster_flags = L.RASTER_PERSPECTIVE_CORRECT if perspective_correct else 0
        if isinstance(mano_path, dict):
            self.mano_layer = MANO_SMPL(mano_path, dataset)
        else:
            self.mano_layer = MANO_SMPL(mano_path + "/MANO_RIGHT.pkl", dataset)
        self.paras = cam_para
        self.img_size = image_size
        self.crop_size = crop_size
        if crop_size[0] != crop_size[1]:
            raise ValueError("square crops only (the reference uses crop_size[0] for both axes)")
        if mode not in ("literal", "direct"):
            raise ValueError("mode must be 'literal' or 'direct'")
        self.mode = mode
        self._intr = (C.c_float * 4)(*[float(v) for v in cam_para])
        if dataset == "nyu":
            self.depth_range = [500, 1200]
        if dataset == "msra" or dataset == "icvl":
            self.depth_range = [150, 600]
        dev = self.mano_layer.v_template.device
        R = crop_size[0]
        g = (2 * (torch.arange(R, dtype=torch.float32) + 0.5) / R - 1.0)
        yy, xx = torch.meshgrid(g, g, indexing="ij")
        self.xy_mesh = torch.stack((xx, yy), -1).reshape(1, -1, 2).to(dev)

    # -- view set-up + rasterisation -----------------------------------------------------------------
    def _view(self, center3d, cube_size, M=None):
        lib = L.lib()
        center3d = L.f32c(center3d)
        cube_size = L.f32c(cube_size)
        B = center3d.shape[0]
        R = self.crop_size[0]
        dev = center3d.device
        view = torch.empty(B, L.VIEW_STRIDE, device=dev)
        xs = torch.empty(B, R, device=dev)
        ys = torch.empty(B, R, device=dev)
        M_out = torch.empty(B, 3, 3, device=dev)
        M_in = None
        if M is not None:
            M_in = L.f32c(M)
            if float(M_in[:, 0, 1].abs().max()) != 0.0 or float(M_in[:, 1, 0].abs().max()) != 0.0:
                raise ValueError("only axis-aligned crop transforms are supported")
        mode = 0 if (self.mode == "direct" and M is None) else 1
        L.check(lib.dsf_view_setup(mode, B, center3d.data_ptr(), cube_size.data_ptr(), self._intr,
                                   int(self.img_size[0]), int(self.img_size[1]), R, L.ptr(M_in),
                                   view.data_ptr(), xs.data_ptr(), ys.data_ptr(), M_out.data_ptr(),
                                   L.stream_ptr()))
        return view, xs, ys, M_out

    def _rasterize(self, hand_verts, center3d, cube_size, M=None):
        view, xs, ys, M_out = self._view(center3d, cube_size, M)
        img, p2f = _RasterFunction.apply(self.mano_layer, hand_verts, view, xs, ys, self.raster_flags)
        return img, p2f, M_out

    @staticmethod
    def _split(model_paras):
        qd = 4 if model_paras.size(-1) == 63 else 3
        return (model_paras[:, :qd], model_paras[:, qd:qd + 45], model_paras[:, qd + 45:qd + 55],
                model_paras[:, qd + 55:])

    # -- R5 entry points -----------------------------------------------------------------------------
    def render(self, model_paras, center3d, cube_size, M=None):
        """mano_layer.py:1071-1097 -> (img, joint_uvd, joint_xyz, mesh_xyz).  The whole chain is one
        autograd node (two C-ABI calls); M is recomputed from center3d / cube_size like the reference
        does (:1088-1089), the argument is accepted and ignored for signature compatibility.
        Gradients flow to model_paras (center3d and cube_size are data)."""
        if model_paras.size(-1) not in (62, 63):
            raise ValueError("model_paras must be (B, 62) or (B, 63)")
        view, xs, ys, M_used = self._view(center3d, cube_size)
        img, joint_uvd, joint_xyz, mesh_xyz, _ = _RenderFunction.apply(self, model_paras, center3d, cube_size, view, xs,
                                                                       ys, M_used)
        return img, joint_uvd, joint_xyz, mesh_xyz

    def render_modular(self, model_paras, center3d, cube_size, M=None):
        """The same computation composed from the individual autograd ops (MANO, raster, torch helpers)."""
        quat, theta, beta, cam = self._split(model_paras)
        verts, joints = self.mano_layer.get_mano_vertices(quat, theta, beta, cam, global_scale=1 / 125)
        hand_verts = verts * cube_size.unsqueeze(1) / 2 + center3d.unsqueeze(1)
        hand_joints = joints * cube_size.unsqueeze(1) / 2 + center3d.unsqueeze(1)
        img, _, M_used = self._rasterize(hand_verts, center3d, cube_size)     # M is recomputed (:1088-1089)
        center2d = self.points3DToImg(center3d.unsqueeze(1)).squeeze(1)
        joint_uvd = self.JointTrans(hand_joints, M_used, center2d, cube_size)
        joint_xyz = (hand_joints - center3d.unsqueeze(1)) / cube_size.unsqueeze(1) * 2
        mesh_xyz = (hand_verts - center3d.unsqueeze(1)) / cube_size.unsqueeze(1) * 2
        return img, joint_uvd, joint_xyz, mesh_xyz

    def normal_render(self, model_paras, center3d, cube_size):
        quat, theta, beta, cam = self._split(model_paras[:, :62])
        verts, joints = self.mano_layer.get_mano_vertices(quat, theta, beta, cam, global_scale=1 / 125)
        hand_verts = (verts + 1) / 2 * cube_size.unsqueeze(1) + center3d.unsqueeze(1)
        hand_joints = (joints + 1) / 2 * cube_size.unsqueeze(1) + center3d.unsqueeze(1)
        img, _, M_used = self._rasterize(hand_verts, center3d, cube_size)
        center2d = self.points3DToImg(center3d.unsqueeze(1)).squeeze(1)
        joint_uvd = self.JointTrans(hand_joints, M_used, center2d, cube_size)
        joint_xyz = (hand_joints - center3d.unsqueeze(1)) / cube_size.unsqueeze(1) * 2
        verts_xyz = (hand_verts - center3d.unsqueeze(1)) / cube_size.unsqueeze(1) * 2
        return img, joint_uvd, joint_xyz, verts_xyz

    def forward(self, model_paras, center3d, cube_size, augmentView=None, augmentShape=None,
                augmentCenter=None, augmentSize=None, mask=True):
        device = model_paras.device
        B = model_paras.size(0)
        quat, theta, beta, cam = self._split(model_paras)
        if augmentShape is not None:
            beta = beta + augmentShape
        hand_verts, hand_joints = self.mano_layer.get_mano_vertices(quat, theta, beta, cam)
        synth_center = hand_joints.mean(dim=1, keepdim=True).clone()
        hand_verts = hand_verts - synth_center
        hand_joints = hand_joints - synth_center
        if center3d is None:
            depth = torch.rand([B, 1]) * (self.depth_range[1] - self.depth_range[0]) + self.depth_range[0]
            center3d = torch.cat((torch.zeros([B, 2]), depth), dim=-1).to(device)
        hand_verts = hand_verts + center3d.unsqueeze(1)
        hand_joints = hand_joints + center3d.unsqueeze(1)
        if augmentView is not None:
            hand_verts, hand_joints = RotationPoints(hand_verts, hand_joints, center3d, augmentView)
        if augmentCenter is not None:
            center3d = center3d + augmentCenter
        if augmentSize is not None:
            cube_size = cube_size * augmentSize
        img, _, M = self._rasterize(hand_verts, center3d, cube_size)
        center2d = self.points3DToImg(center3d.unsqueeze(1)).squeeze(1)
        joint_uvd = self.JointTrans(hand_joints, M, center2d, cube_size)
        verts_uvd = self.JointTrans(hand_verts, M, center2d, cube_size)
        joint_xyz = (hand_joints - center3d.unsqueeze(1)) / cube_size.unsqueeze(1) * 2
        verts_xyz = (hand_verts - center3d.unsqueeze(1)) / cube_size.unsqueeze(1) * 2
        if mask:
            img = self.mask_img(img, joint_uvd, 0.15, 0.3)
        return img, joint_uvd, verts_uvd, joint_xyz, verts_xyz, center3d, cube_size, M

    def M_render(self, model_paras, center3d, cube_size, M=None, mask=True):
        quat, theta, beta, cam = self._split(model_paras)
        hand_verts, hand_joints = self.mano_layer.get_mano_vertices(quat, theta, beta, cam)
        img, _, M_used = self._rasterize(hand_verts, center3d, cube_size, M)
        if mask:
            center2d = self.points3DToImg(center3d.unsqueeze(1)).squeeze(1)
            joint_uvd = self.JointTrans(hand_joints, M_used, center2d, cube_size)
            img = self.mask_img(img, joint_uvd, 0.15, 0.3)
        return img

    def get_mesh_xyz(self, model_paras):
        quat, theta, beta, cam = self._split(model_paras[:, :62])
        hand_mesh, hand_joints = self.mano_layer.get_mano_vertices(quat, theta, beta, cam, global_scale=1 / 125)
        return hand_joints, hand_mesh

    def get_mesh_xyz_old(self, model_paras):
        hand_joints, hand_mesh = self.get_mesh_xyz(model_paras)
        return hand_joints + 1, hand_mesh + 1

    def mesh2img(self, hand_mesh, center3d, cube_size):
        img, _, _ = self._rasterize(hand_mesh, center3d, cube_size)
        return img

    def getDepth(self, hand_verts, hand_joints, center3d, cube_size, M, rot=None):
        if rot is not None:
            hand_verts, hand_joints = RotationPoints(hand_verts, hand_joints, center3d, rot)
        img, _, M_used = self._rasterize(hand_verts, center3d, cube_size, M)
        center2d = self.points3DToImg(center3d.unsqueeze(1)).squeeze(1)
        return img, self.JointTrans(hand_joints, M_used, center2d, cube_size)

    # -- small camera helpers (cheap elementwise torch, same formulas as :1133-1169, :1289-1324) -----
    def comToBounds(self, com, size):
        fx, fy, _, _ = self.paras
        zstart = com[:, 2] - size[:, 2] / 2.
        zend = com[:, 2] + size[:, 2] / 2.
        ax = com[:, 0] * com[:, 2] / fx
        ay = com[:, 1] * com[:, 2] / fy
        xstart = torch.floor((ax - size[:, 0] / 2.) / com[:, 2] * fx + 0.5).int()
        xend = torch.floor((ax + size[:, 0] / 2.) / com[:, 2] * fx + 0.5).int()
        ystart = torch.floor((ay - size[:, 1] / 2.) / com[:, 2] * fy + 0.5).int()
        yend = torch.floor((ay + size[:, 1] / 2.) / com[:, 2] * fy + 0.5).int()
        return xstart, xend, ystart, yend, zstart, zend

    def Offset2Trans(self, xstart, xend, ystart, yend):
        R = self.crop_size[0]
        wb, hb = xend - xstart, yend - ystart
        wide = wb > hb
        sz0 = torch.where(wide, torch.full_like(wb, R), (wb * R / hb).int())
        sz1 = torch.where(wide, (hb * R / wb).int(), torch.full_like(wb, R))
        s = torch.where(wide, R / wb, R / hb)
        ox = torch.floor(R / 2. - sz0 / 2.)
        oy = torch.floor(R / 2. - sz1 / 2.)
        M = torch.zeros(xstart.size(0), 3, 3, device=xstart.device)
        M[:, 0, 0] = s
        M[:, 1, 1] = s
        M[:, 2, 2] = 1
        M[:, 0, 2] = ox - s * xstart
        M[:, 1, 2] = oy - s * ystart
        return M

    def normalize_img(self, imgD, com, cube):
        z_min = (com[:, 2] - cube[:, 2] / 2.).view(-1, 1, 1, 1)
        z_max = (com[:, 2] + cube[:, 2] / 2.).view(-1, 1, 1, 1)
        imgD = torch.where((imgD == -1) | (imgD == 0), z_max, imgD)
        imgD = torch.minimum(torch.maximum(imgD, z_min), z_max)
        return (imgD - com[:, 2].view(-1, 1, 1, 1)) / (cube[:, 2].view(-1, 1, 1, 1) / 2.)

    def JointTrans(self, joint, M, com, cube):
        uvd = self.points3DToImg(joint)
        ones = torch.ones_like(uvd[:, :, :1])
        uv = torch.matmul(M.unsqueeze(1), torch.cat((uvd[:, :, 0:2], ones), dim=-1).unsqueeze(-1)).squeeze(-1)
        uv = uv[:, :, 0:2] / self.crop_size[0] * 2 - 1
        d = (uvd[:, :, 2:] - com.unsqueeze(1)[:, :, 2:]) / (cube.unsqueeze(1)[:, :, 2:] / 2.0)
        return torch.cat((uv, d), dim=-1)

    def pointsImgTo3D(self, point_uvd):
        fx, fy, fu, fv = self.paras
        x = (point_uvd[:, :, 0] - fu) * point_uvd[:, :, 2] / fx
        y = (point_uvd[:, :, 1] - fv) * point_uvd[:, :, 2] / fy
        return torch.stack((x, y, point_uvd[:, :, 2]), dim=-1)

    def points3DToImg(self, joint_xyz):
        fx, fy, fu, fv = self.paras
        u = joint_xyz[:, :, 0] * fx / (joint_xyz[:, :, 2] + 1e-8) + fu
        v = joint_xyz[:, :, 1] * fy / (joint_xyz[:, :, 2]) + fv
        return torch.stack((u, v, joint_xyz[:, :, 2]), dim=-1)

    def mask_img(self, img, img_joint, mask_offset, mask_para, min_mask_num=3, max_mask_num=10):
        """Random spherical occluders around a few joints (non-differentiable augmentation, :1326-1340).
        The random numbers are drawn with the reference's own calls in its order (numpy choice twice, torch.rand
        on the CPU twice), so equal seeds give equal occluders; the masking itself is one kernel (dsf_mask_img)
        instead of a (B, mask_num, R*R) distance tensor."""
        device = img.device
        b, j, _ = img_joint.size()
        mask_num = int(np.random.choice(np.arange(min_mask_num, max_mask_num), 1, replace=False)[0])
        joint_id = np.random.choice(np.arange(0, j), mask_num, replace=False)
        centre = img_joint[:, joint_id, :] + ((torch.rand(b, mask_num, 3) - 0.5) * mask_offset * 2).to(device)
        radius = torch.rand([b, mask_num]).to(device) * mask_para
        return self.mask_spheres(img, centre, radius)

    def mask_spheres(self, img, centres, radii):
        """The deterministic part of mask_img: pixels inside any sphere (centres (B,n,3), radii (B,n)) -> 1.0."""
        img_c = L.f32c(img.detach())
        R = img_c.shape[-1]
        if img_c.shape[-2] != R:
            raise ValueError("square images only")
        B = img_c.numel() // (R * R)
        centres, radii = L.f32c(centres.detach()), L.f32c(radii.detach())
        out = torch.empty_like(img_c)
        L.check(L.lib().dsf_mask_img(B, R, img_c.data_ptr(), int(radii.shape[1]), centres.data_ptr(), radii.data_ptr(),
                                     out.data_ptr(), L.stream_ptr()))
        return out

    def synth2real(self, noraml_img, noise=0.1, noise_patch=2, sigma=1.7, bk_value=0.95):
        """:1222-1231 - patch-wise white noise on the foreground, then a 5x5 Gaussian on the reflect-padded image
        (one kernel, dsf_synth2real).  The reference reads ``self.smoothing``, which it never assigns; the
        filter is its own ``GaussianSmoothing(5)`` (:808-868), the only reading under which the 2-pixel padding
        gives back an image of the same size.  The noise is drawn like the reference does (torch.randn on the
        CPU), so equal seeds give equal images."""
        B, C_, H, W = noraml_img.size()
        if C_ != 1 or H != W:
            raise ValueError("expected (B,1,R,R)")
        img_c = L.f32c(noraml_img.detach())
        nz = L.f32c(noise * torch.randn((B, C_, H // noise_patch, W // noise_patch)))
        out = torch.empty_like(img_c)
        L.check(L.lib().dsf_synth2real(B, H, img_c.data_ptr(), nz.data_ptr(), int(noise_patch), float(bk_value),
                                       float(sigma), out.data_ptr(), L.stream_ptr()))
        return out

    # -- resampling helpers of the reference's own pixel chain (:1233-1287); Render itself never materialises
    #    the 640^2 raster, these exist for callers that hold full-size images -------------------------
    def resize(self, img):
        """:1233-1242 - nearest resample of (B,C,S,S) to the sensor size (H,W)."""
        import torch.nn.functional as F
        b = img.size(0)
        theta = torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], device=img.device).unsqueeze(0).repeat(b, 1, 1)
        grid = F.affine_grid(theta, [b, 1, int(self.img_size[1]), int(self.img_size[0])], align_corners=False)
        return F.grid_sample(img, grid, mode="nearest", align_corners=False)

    def affine_grid(self, img, M):
        """:1244-1255 - sampling grid of the crop: crop pixel (cx, cy) reads sensor position M^-1 (cx, cy, 1)."""
        b, _, h_ori, w_ori = img.size()
        R = self.crop_size[0]
        ii = torch.arange(R, device=img.device, dtype=torch.float32)
        yy, xx = torch.meshgrid(ii, ii, indexing="ij")
        mesh = torch.stack((xx, yy, torch.ones_like(xx)), -1).reshape(1, -1, 3, 1)
        pts = torch.matmul(torch.inverse(M).view(b, 1, 3, 3), mesh).squeeze(-1)[:, :, 0:2]
        scale = torch.tensor([w_ori, h_ori], device=img.device, dtype=torch.float32).view(1, 1, 2)
        return (pts / scale * 2 - 1).view(b, R, R, 2)

    def warpPerspective(self, img, M):
        """:1257-1260 - nearest crop of a sensor-size image through M."""
        import torch.nn.functional as F
        return F.grid_sample(img, self.affine_grid(img, M), mode="nearest", align_corners=False)

    def ResizeRenderImg(self, img):
        """:1262-1273 - RoIAlign (sampling_ratio 1, one bilinear sample per output pixel) of the square raster to
        the sensor size: output pixel (r, c) samples ((c + .5) S / W - .5, (r + .5) S / H - .5)."""
        import torch.nn.functional as F
        b = img.size(0)
        S = float(max(self.img_size))
        W, H = int(self.img_size[0]), int(self.img_size[1])
        xs = (torch.arange(W, device=img.device, dtype=torch.float32) + 0.5) * (S / W) - 0.5
        ys = (torch.arange(H, device=img.device, dtype=torch.float32) + 0.5) * (S / H) - 0.5
        gx = (xs + 0.5) / img.size(-1) * 2 - 1
        gy = (ys + 0.5) / img.size(-2) * 2 - 1
        grid = torch.stack((gx.view(1, W).expand(H, W), gy.view(H, 1).expand(H, W)), -1).unsqueeze(0).repeat(b, 1, 1, 1)
        return F.grid_sample(img, grid, mode="bilinear", padding_mode="border", align_corners=False)

    def massCenter(self, img):
        """:1275-1287 - (u, v, depth) centroid of the positive pixels of (B,1,H,W)."""
        b, _, h, w = img.size()
        yv, xv = torch.meshgrid(torch.arange(h, device=img.device, dtype=torch.float32),
                                torch.arange(w, device=img.device, dtype=torch.float32), indexing="ij")
        fg = img.gt(0).float()
        pts = torch.cat((xv.expand(b, 1, h, w), yv.expand(b, 1, h, w), img), dim=1) * fg
        return pts.mean(-1).mean(-1) / fg.mean(-1).mean(-1)
