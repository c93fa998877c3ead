"""The rarely used variants of the sphere model (render_model/mano_layer.py:319-372, :388-567): importable,
with real bodies, behind the reference's names.  None of them is called by the trainer ("live
vs dead API surface"), so they are kept in plain torch on top of the sphere-set kernel (dsf_sphere_set) instead
of getting kernels of their own.  Centres are differentiable (they are an interpolation of the joints); radii
come from the kernel and carry no gradient."""
from __future__ import annotations

import numpy as np
import torch

# finger sphere (3 per bone x 15 bones, tips folded in) -> 21-joint label / finger label  (:476, :527)
_ID21 = [1, 1, 2, 2, 2, 3, 3, 3, 16, 4, 4, 5, 5, 5, 6, 6, 6, 17, 7, 7, 8, 8, 8, 9, 9, 9, 18,
         10, 10, 11, 11, 11, 12, 12, 12, 19, 13, 13, 14, 14, 14, 15, 15, 15, 20]
_ID5 = [k // 9 + 1 for k in range(45)]
# finger spheres owned by each of the 20 non-palm joints (:497-502): per finger 2 + 3 + 3 along the chain, the 9th is the tip
_GROUP21 = [[9 * f + o for o in offs] for f in range(5) for offs in ([0, 1], [2, 3, 4], [5, 6, 7])] + \
           [[9 * f + 8] for f in range(5)]
N_PALM = 21


class SphereVariants:
    """mixin of MANO_SMPL; needs self._sphere_set, self.child, self.interval, self.vertex_joint_index_list"""

    # -- pieces of get_sphere_radius under their own names (:319-372) ---------------------------------
    def get_sphere(self, joints):
        """(B,21,3) -> (B,66,3): palm root + 4 interpolated spheres per metacarpal, then 3 per finger bone."""
        B = joints.shape[0]
        t_palm = torch.arange(1, 5, device=joints.device, dtype=joints.dtype) / 5.0          # plam_interval_value
        t_fing = torch.arange(0, 3, device=joints.device, dtype=joints.dtype) / 3.0          # interval_value
        root = joints[:, 0:1]
        palm = root[:, :, None] + (joints[:, [1, 4, 7, 10, 13]] - root)[:, :, None] * t_palm.view(1, 1, -1, 1)
        parent = joints[:, 1:16]
        fing = parent[:, :, None] + (joints[:, self.child] - parent)[:, :, None] * t_fing.view(1, 1, -1, 1)
        return torch.cat((root, palm.reshape(B, -1, 3), fing.reshape(B, -1, 3)), dim=1)

    def get_radius(self, joints, mesh):
        """(B,21,3), (B,779,3) -> (B,66) sphere radii (forward only)."""
        return self._sphere_set(joints, joints, mesh)[1]

    def calculate_PWE_coll(self, joints_PWE, joints, meshs):
        """:388-400 - the collision hinge with centres from ``joints_PWE`` and radii from ``joints`` / ``meshs``."""
        c = self.get_sphere(joints_PWE)
        r = self.get_radius(joints, meshs)
        d = c[:, :, None] - c[:, None]
        dis = torch.sqrt((d * d).sum(-1) + 1e-8)
        err = torch.clamp(r[:, :, None] + r[:, None] - dis, min=0) * self.mask.to(c.device)
        # the reference sums a size-1 axis the second time (:398): the 0.1 gate acts per sphere row
        gate = err.sum(-1, keepdim=True).sum(-1, keepdim=True).lt(0.1).float()
        return torch.mean((err * gate).sum(-1))

    # -- segmentation variants (:462-488, :513-539) ---------------------------------------------------
    def _seg_by_spheres(self, joints, joints_mano, mesh, pcl, id_map):
        c, _ = self._sphere_set(joints, joints, mesh)
        _, r = self._sphere_set(joints_mano, joints_mano, mesh)
        pcl = pcl.to(c.device).float()

        def surf(cs, rs):
            d = torch.sqrt(((pcl.unsqueeze(-2) - cs.unsqueeze(1)) ** 2).sum(-1) + 1e-8)
            return torch.abs(d - rs.unsqueeze(1)).min(-1)

        fd, fi = surf(c[:, N_PALM:], r[:, N_PALM:])
        pd, _ = surf(c[:, :N_PALM], r[:, :N_PALM])
        lab = torch.as_tensor(id_map, device=c.device)[fi]
        return torch.where(pd < fd, torch.zeros_like(lab), lab)

    def seg_pcl_21(self, joints, joints_mano, mesh, pcl):
        return self._seg_by_spheres(joints, joints_mano, mesh, pcl, _ID21)

    def seg_pcl_finger(self, joints, joints_mano, mesh, pcl):
        return self._seg_by_spheres(joints, joints_mano, mesh, pcl, _ID5)

    # -- point -> sphere / vertex distances per part (:429-460, :490-511, :541-567) -------------------
    def _point2sphere(self, joint, mesh, pcl, pcl_seg, groups, first_label):
        c, r = self._sphere_set(joint, joint, mesh)
        pcl = pcl.to(c.device).float()
        out, min_index = [], torch.zeros_like(pcl_seg)
        for k, g in enumerate(groups):
            idx = torch.as_tensor(g, device=c.device) + N_PALM
            d = torch.sqrt(((pcl.unsqueeze(-2) - c[:, idx].unsqueeze(1)) ** 2).sum(-1) + 1e-8)
            d = torch.abs(d - r[:, idx].unsqueeze(1))
            own = pcl_seg.eq(k + first_label)
            d, arg = torch.where(own.unsqueeze(-1), d, torch.zeros_like(d)).min(-1)
            n = d.gt(0).sum(-1)
            loss = d.sum(-1) / (n + 1e-8)
            out.append(torch.where(n.eq(0), torch.zeros_like(loss), loss))
            min_index = torch.where(own, idx[arg], min_index)
        return torch.stack(out, dim=-1), min_index

    def calculate_point2shpere_distance(self, joint, mesh, pcl, pcl_seg):
        groups = [list(range(self.interval * b, self.interval * (b + 1))) for b in range(15)]
        return self._point2sphere(joint, mesh, pcl, pcl_seg, groups, 1)

    def calculate_point2shpere_distance_21(self, joint, mesh, pcl, pcl_seg):
        return self._point2sphere(joint, mesh, pcl, pcl_seg, _GROUP21, 1)[0]

    def calculate_point2shpere_distance_finger(self, joint, mesh, pcl, pcl_seg):
        groups = [list(range(9 * f, 9 * (f + 1))) for f in range(5)]
        return self._point2sphere(joint, mesh, pcl, pcl_seg, groups, 1)[0]

    def calculate_point2mesh_distance(self, mesh, pcl, pcl_seg):
        """:429-441 - per part k (vertices whose largest skin weight is joint k): mean over the cloud of the
        squared distance to the nearest such vertex, points of other labels counting 1e5."""
        verts = mesh[:, :-1]
        out = []
        for k in range(15):
            sel = verts[:, self.vertex_joint_index_list[k].to(verts.device)]
            d = ((pcl.unsqueeze(-2) - sel.unsqueeze(1)) ** 2).sum(-1)
            d = torch.where(pcl_seg.eq(k).unsqueeze(-1), d, torch.full_like(d, 1e5))
            out.append(d.min(-1)[0].mean(-1))
        return torch.stack(out, dim=-1)
