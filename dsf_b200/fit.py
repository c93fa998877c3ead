"""Fused model-fitting step (dsf_fit_step): MANO forward -> rasterise -> m2d depth loss -> backward
to the 62 MANO/camera parameters, with persistent buffers and optional CUDA-graph replay.

This is the unit bench.py times: one call = one pass of the hot path over one batch of hands.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


# sampling modes of dsf_view_setup: "direct" = R x R raster sampled at crop pixel centres (benchmark configs),
# "literal" = the reference's 640^2 -> resize -> crop chain, "direct_aligned" = direct raster whose sample i sits at
# crop coordinate i, i.e. registered with M / JointTrans / loader-cropped depth (no half-pixel offset)
_MODES = {"direct": 0, "literal": 1, "direct_aligned": 2}


class FitStep:
    """One fitting step over a resident batch.

    perspective_correct: pytorch3d's RasterizationSettings flag; False (default) is what the reference runs
    with (pytorch3d 0.4.0 default, see include/dsf_b200.h).  keep_pix_to_face=False drops the pix_to_face
    plane, which nothing reads once the rasteriser emits the vertex gradient itself (default settings only).
    The per-hand view records depend only on (center3d, cube): they are rebuilt by set_inputs(), not by
    every step().  fuse_target_rows=True: a row-run packed target (pcl.RowRunTarget) handed to set_inputs() is
    not unpacked into ``self.target``; the rasteriser's epilogue decodes it in place (dsf_fit_step_rows - no unpack
    launch, no fp32 target plane; ``materialise_target()`` rebuilds ``self.target`` on demand)."""

    def __init__(self, mano_layer, batch, crop=128, cam_para=(588.03, 587.07, 320.0, 240.0),
                 image_size=(640, 480), mode="direct", loss_weight=0.1, use_graph=True, device=None, chunks=1,
                 perspective_correct=False, keep_pix_to_face=True, fuse_target_rows=False):
        self.lib = L.lib()
        self.flags = L.RASTER_PERSPECTIVE_CORRECT if perspective_correct else 0
        if mode == "literal":
            # literal 640-pixel raster: float32 sample coordinates are not exactly 1 - (2 q + 1) / S there, keep
            # the per-pixel backward kernel (see DSF_RASTER_SEPARATE_BACKWARD)
            self.flags |= L.RASTER_SEPARATE_BACKWARD
        if self.flags and not keep_pix_to_face:
            raise ValueError("only the direct-mode, non-perspective-correct step can drop the pix_to_face plane")
        if self.flags and fuse_target_rows:
            raise ValueError("only the direct-mode, non-perspective-correct step can decode the row-run target in place")
        self.fuse_target_rows = bool(fuse_target_rows)
        self._rows_active = False         # the step reads the row-run buffers instead of self.target
        self.layer = mano_layer
        self.B, self.R = int(batch), int(crop)
        self.mode = _MODES[mode]
        self.loss_weight = float(loss_weight)
        self.dev = device or mano_layer.v_template.device
        self.W, self.H = int(image_size[0]), int(image_size[1])
        self._intr = (C.c_float * 4)(*[float(v) for v in cam_para])
        B, R, dev = self.B, self.R, self.dev
        f = lambda *s: torch.empty(*s, device=dev)
        self.params = f(B, 62)
        self.center3d = f(B, 3)
        self.cube = f(B, 3)
        self.target = f(B, R, R)
        self.target_u16 = None            # allocated on first use (set_inputs with a uint16 target)
        self._row_bufs = None             # ... with a row-run packed target
        self.view = f(B, L.VIEW_STRIDE)
        self.xs = f(B, R)
        self.ys = f(B, R)
        self.M = f(B, 3, 3)
        self.img = f(B, R, R)
        self.p2f = torch.empty(B, R, R, dtype=torch.int32, device=dev) if keep_pix_to_face else None
        self.verts = f(B, L.NVW, 3)
        self.joints = f(B, L.NJOUT, 3)
        self.g_params = f(B, 62)
        self.parts = f(B, 2)
        self.totals = f(4)
        # the batch can be cut into `chunks` slices that run the kernel chain on parallel streams
        # (hands are independent): the tail of one slice's kernels overlaps the next slice's heads
        self.chunks = max(1, min(int(chunks), B))
        base, rem = divmod(B, self.chunks)
        self._bounds, lo = [], 0
        for c in range(self.chunks):
            hi = lo + base + (1 if c < rem else 0)
            self._bounds.append((lo, hi))
            lo = hi
        self.ws = [f(self.lib.dsf_fit_workspace_floats(hi - lo, R)) for lo, hi in self._bounds]
        self.chunk_totals = f(self.chunks, 4)
        self._streams = [torch.cuda.Stream(device=dev) for _ in range(self.chunks - 1)]
        self.crop_joints = None          # (B,J,3) teacher joints -> crop_hand before the loss
        self.use_graph = use_graph
        self._graph = None
        self.launches_per_step = 0
        self.setup_launches = 0

    # inputs already resident in HBM -----------------------------------------------------------------
    def _view_setup(self):
        L.check(self.lib.dsf_view_setup(self.mode, self.B, self.center3d.data_ptr(), self.cube.data_ptr(),
                                        self._intr, self.W, self.H, self.R, None, self.view.data_ptr(),
                                        self.xs.data_ptr(), self.ys.data_ptr(), self.M.data_ptr(), L.stream_ptr()))
        self.setup_launches = self.lib.dsf_last_launch_count()

    def set_inputs(self, params, center3d, cube, target=None):
        self.params.copy_(params, non_blocking=True)
        self.center3d.copy_(center3d, non_blocking=True)
        self.cube.copy_(cube, non_blocking=True)
        self._view_setup()                       # view records / sample grids follow (center3d, cube)
        if target is not None:
            from .pcl import RowRunTarget, target_from_u16_rows
            if isinstance(target, RowRunTarget):
                # row-run packed sensor crop: only the non-background span of every row travels over PCIe
                if self._row_bufs is None:
                    dev = self.target.device
                    self._row_bufs = (torch.empty(self.B, self.R, 2, dtype=torch.uint16, device=dev),
                                      torch.empty(self.B + 1, dtype=torch.int32, device=dev),
                                      torch.empty(self.B * self.R * self.R, dtype=torch.uint16, device=dev))
                if self.fuse_target_rows:
                    d_rows, d_off, d_pay = self._row_bufs
                    d_rows.copy_(target.rows, non_blocking=True)
                    d_off.copy_(target.hand_offset, non_blocking=True)
                    d_pay[:target.payload.numel()].copy_(target.payload, non_blocking=True)
                    self._set_rows_active(True)
                    return
                target_from_u16_rows(target, self.center3d, self.cube, 0, out=self.target, buffers=self._row_bufs)
            elif target.dtype == torch.uint16:
                # sensor format: uint16 millimetres travel over PCIe (half the bytes), normalised here
                if self.target_u16 is None:
                    self.target_u16 = torch.empty(self.B, self.R, self.R, dtype=torch.uint16, device=self.target.device)
                self.target_u16.copy_(target.reshape(self.B, self.R, self.R), non_blocking=True)
                L.check(self.lib.dsf_target_from_u16(self.B, self.R, self.target_u16.data_ptr(),
                                                     self.center3d.data_ptr(), self.cube.data_ptr(), 0,
                                                     self.target.data_ptr(), L.stream_ptr()))
            else:
                self.target.copy_(target.reshape(self.B, self.R, self.R), non_blocking=True)
            self._set_rows_active(False)

    def _set_rows_active(self, on):
        if on != self._rows_active:
            self._rows_active = on
            self._graph = None               # the captured launch sequence reads the other target buffers

    def materialise_target(self):
        """Rebuild ``self.target`` (normalised fp32) from the row-run buffers of the last set_inputs()."""
        if self._row_bufs is None:
            return self.target
        d_rows, d_off, d_pay = self._row_bufs
        L.check(self.lib.dsf_target_from_u16_rows(self.B, self.R, d_rows.data_ptr(), d_off.data_ptr(), d_pay.data_ptr(),
                                                  self.center3d.data_ptr(), self.cube.data_ptr(), 0,
                                                  self.target.data_ptr(), L.stream_ptr()))
        return self.target

    def set_crop_joints(self, joints):
        """Teacher joints (B,J,3) in normalised cube units: the rendered image is passed through
        crop_hand (data/render_loader.py:1209) before the m2d loss, as train_render.py:727 does.
        The teacher joints change with every batch (train_render.py:727): later calls with the same shape
        copy into the buffer the captured graph reads; only enabling cropping, or changing the joint count,
        after capture is refused."""
        joints = L.f32c(joints)
        if self.crop_joints is not None and tuple(self.crop_joints.shape) == tuple(joints.shape):
            self.crop_joints.copy_(joints, non_blocking=True)
            return
        if self._graph is not None:
            raise RuntimeError("crop joints must first be set (and keep their shape) before the CUDA graph is captured")
        self.crop_joints = joints.clone()

    def _enqueue(self):
        n = 0
        nj = 0 if self.crop_joints is None else self.crop_joints.shape[1]
        R2 = self.R * self.R

        def run_chunk(c):
            lo, hi = self._bounds[c]
            off = lambda t, per: t.data_ptr() + lo * per * t.element_size()
            if self._rows_active:
                d_rows, d_off, d_pay = self._row_bufs
                L.check(self.lib.dsf_fit_step_rows(
                    self.layer._handle, hi - lo, self.R, off(self.params, 62), off(self.center3d, 3),
                    off(self.cube, 3), off(self.view, L.VIEW_STRIDE), off(self.xs, self.R), off(self.ys, self.R),
                    off(d_rows, 2 * self.R), off(d_off, 1), d_pay.data_ptr(), 0, self.loss_weight, self.B,
                    None if self.crop_joints is None else off(self.crop_joints, nj * 3), nj, off(self.M, 9),
                    self._intr, off(self.img, R2), None if self.p2f is None else off(self.p2f, R2),
                    off(self.verts, L.NVW * 3), off(self.joints, L.NJOUT * 3), off(self.g_params, 62), off(self.parts, 2),
                    self.totals.data_ptr() if self.chunks == 1 else self.chunk_totals[c].data_ptr(),
                    self.ws[c].data_ptr(), self.flags, L.stream_ptr()))
                return self.lib.dsf_last_launch_count()
            L.check(self.lib.dsf_fit_step(
                self.layer._handle, hi - lo, self.R, off(self.params, 62), off(self.center3d, 3),
                off(self.cube, 3), off(self.view, L.VIEW_STRIDE), off(self.xs, self.R), off(self.ys, self.R),
                off(self.target, R2), self.loss_weight, self.B,
                None if self.crop_joints is None else off(self.crop_joints, nj * 3), nj, off(self.M, 9),
                self._intr, off(self.img, R2), None if self.p2f is None else off(self.p2f, R2),
                off(self.verts, L.NVW * 3), off(self.joints, L.NJOUT * 3), off(self.g_params, 62), off(self.parts, 2),
                self.totals.data_ptr() if self.chunks == 1 else self.chunk_totals[c].data_ptr(),
                self.ws[c].data_ptr(), self.flags, L.stream_ptr()))
            return self.lib.dsf_last_launch_count()

        main = torch.cuda.current_stream()
        for c, st in enumerate(self._streams, start=1):        # fork
            st.wait_stream(main)
            with torch.cuda.stream(st):
                n += run_chunk(c)
        n += run_chunk(0)
        for st in self._streams:                                # join
            main.wait_stream(st)
        if self.chunks > 1:
            # totals of the whole batch from the per-slice records ([3] is the un-normalised loss)
            L.check(self.lib.dsf_sum_totals(self.chunks, self.chunk_totals.data_ptr(), self.B,
                                            self.totals.data_ptr(), L.stream_ptr()))
            n += self.lib.dsf_last_launch_count()
        self.launches_per_step = n

    def step(self):
        """Enqueue one fitting step on the current stream; results stay on the device
        (self.totals[0] = loss, self.g_params = d loss / d params, self.img = rendered depth)."""
        if not self.use_graph:
            self._enqueue()
            return
        if self._graph is None:
            self._enqueue()                      # warm-up outside capture (lazy module load, attributes)
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self._graph = g
        self._graph.replay()

    def render_target(self, params_target):
        """Fill self.target with the rendering of another parameter set (synthetic 'real' depth)."""
        keep = self.params.clone()
        self.params.copy_(params_target)
        self._set_rows_active(False)
        self.target.fill_(1.0)
        self._view_setup()
        self._enqueue()
        self.target.copy_(self.img)
        self.params.copy_(keep)


class MultiViewFitStep:
    """Fused multi-view fitting step (dsf_fit_step_views): MANO once per hand, V rigidly rotated views per
    hand (RotationPoints about center3d, the reference's `getDepth(..., rot)` pattern), one depth +
    silhouette (union-mask m2d) loss over all B*V images and the summed adjoint back to the 62 parameters.
    Buffers are persistent; the launch sequence is replayed as a CUDA graph."""

    def __init__(self, mano_layer, batch, views, crop=256, cam_para=(588.03, 587.07, 320.0, 240.0),
                 image_size=(640, 480), mode="direct", loss_weight=0.1, use_graph=True, device=None,
                 perspective_correct=False, keep_pix_to_face=True):
        self.lib = L.lib()
        self.flags = L.RASTER_PERSPECTIVE_CORRECT if perspective_correct else 0
        if mode == "literal":
            self.flags |= L.RASTER_SEPARATE_BACKWARD
        if self.flags and not keep_pix_to_face:
            raise ValueError("only the direct-mode, non-perspective-correct step can drop the pix_to_face plane")
        self.layer = mano_layer
        self.B, self.V, self.R = int(batch), int(views), int(crop)
        self.mode = _MODES[mode]
        self.loss_weight = float(loss_weight)
        self.dev = device or mano_layer.v_template.device
        self.W, self.H = int(image_size[0]), int(image_size[1])
        self._intr = (C.c_float * 4)(*[float(v) for v in cam_para])
        B, V, R, dev = self.B, self.V, self.R, self.dev
        f = lambda *s: torch.empty(*s, device=dev)
        self.params, self.center3d, self.cube = f(B, 62), f(B, 3), f(B, 3)
        self.rot = f(B, V, 3, 3)
        self.target = f(B * V, R, R)
        self.view, self.xs, self.ys, self.M = f(B * V, L.VIEW_STRIDE), f(B * V, R), f(B * V, R), f(B * V, 3, 3)
        self._c3v, self._cubev = f(B * V, 3), f(B * V, 3)
        self.img = f(B * V, R, R)
        self.p2f = torch.empty(B * V, R, R, dtype=torch.int32, device=dev) if keep_pix_to_face else None
        self.verts, self.joints = f(B, L.NVW, 3), f(B, L.NJOUT, 3)
        self.g_params = f(B, 62)
        self.parts, self.totals = f(B * V, 2), f(4)
        self.ws = f(self.lib.dsf_fit_views_workspace_floats(B, V, R))
        self.use_graph = use_graph
        self._graph = None
        self.launches_per_step = 0

    def set_inputs(self, params, center3d, cube, rot, target=None):
        """rot: (B,V,3) axis-angle or (B,V,3,3) rotation matrices of the views about center3d."""
        from .mano_layer import batch_rodrigues

        self.params.copy_(params, non_blocking=True)
        self.center3d.copy_(center3d, non_blocking=True)
        self.cube.copy_(cube, non_blocking=True)
        rot = rot.to(self.dev)
        if rot.dim() == 3:
            rot = batch_rodrigues(rot.reshape(-1, 3)).reshape(self.B, self.V, 3, 3)
        self.rot.copy_(rot)
        self._c3v.copy_(self.center3d.repeat_interleave(self.V, 0))
        self._cubev.copy_(self.cube.repeat_interleave(self.V, 0))
        self._view_setup()
        if target is not None:
            self.target.copy_(target.reshape(self.B * self.V, self.R, self.R), non_blocking=True)

    def _view_setup(self):
        L.check(self.lib.dsf_view_setup(self.mode, self.B * self.V, self._c3v.data_ptr(), self._cubev.data_ptr(),
                                        self._intr, self.W, self.H, self.R, None, self.view.data_ptr(),
                                        self.xs.data_ptr(), self.ys.data_ptr(), self.M.data_ptr(), L.stream_ptr()))

    def _enqueue(self):
        s = L.stream_ptr()
        L.check(self.lib.dsf_fit_step_views(
            self.layer._handle, self.B, self.V, self.R, self.params.data_ptr(), self.center3d.data_ptr(),
            self.cube.data_ptr(), self.rot.data_ptr(), self.view.data_ptr(), self.xs.data_ptr(), self.ys.data_ptr(),
            self.target.data_ptr(), self.loss_weight, self.img.data_ptr(), L.ptr(self.p2f), self.verts.data_ptr(),
            self.joints.data_ptr(), self.g_params.data_ptr(), self.parts.data_ptr(), self.totals.data_ptr(),
            self.ws.data_ptr(), self.flags, s))
        self.launches_per_step = self.lib.dsf_last_launch_count()

    @property
    def verts_cam(self):
        """(B*V,779,3) camera-space vertices of every view, as the last step left them in the workspace
        (layout of dsf_fit_step_views: MANO scratch | g_verts | verts_cam | ...)."""
        off = self.lib.dsf_mano_workspace_floats(self.B) + self.B * L.NVW * 3
        n = self.B * self.V * L.NVW * 3
        return self.ws[off:off + n].view(self.B * self.V, L.NVW, 3)

    def step(self):
        if not self.use_graph:
            self._enqueue()
            return
        if self._graph is None:
            self._enqueue()
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self._graph = g
        self._graph.replay()

    def render_target(self, params_target):
        keep = self.params.clone()
        self.params.copy_(params_target)
        self.target.fill_(1.0)
        self._enqueue()
        self.target.copy_(self.img)
        self.params.copy_(keep)
