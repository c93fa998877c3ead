"""Intersection-volume metric behind the reference's function names:
``self_intersection`` (eval_coll.py:611-626, over the watertight hand parts of ``get_part_mesh``
:348-373) and ``intersect_vox`` (util/intersect.py:102-107).  The reference runs trimesh on the CPU
(hours for an evaluation set, eval_coll.py:641-674); here one call handles a batch on the GPU."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib as L

PART_PARENT = [0, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13]      # eval_coll.py:615


class PartTopology:
    """Watertight parts over one vertex array: cap loops (centre = mean of the loop, appended after the
    mesh vertices like get_part_mesh does), per-part triangle lists, and the (s, t) pairs to evaluate."""

    def __init__(self, n_verts, cap_loops, part_faces, pair_mask):
        self.n_verts = int(n_verts)
        self.cap_loops = [np.asarray(l, np.int32) for l in cap_loops]
        self.part_faces = [np.asarray(f, np.int32).reshape(-1, 3) for f in part_faces]
        self.n_parts = len(self.part_faces)
        self.pair_mask = np.asarray(pair_mask, np.uint8).reshape(self.n_parts, self.n_parts)
        self.cap_ptr = np.concatenate([[0], np.cumsum([len(l) for l in self.cap_loops])]).astype(np.int32)
        self.cap_idx = (np.concatenate(self.cap_loops) if self.cap_loops else np.zeros(0)).astype(np.int32)
        self.part_ptr = np.concatenate([[0], np.cumsum([len(f) for f in self.part_faces])]).astype(np.int32)
        self.faces = np.concatenate(self.part_faces).astype(np.int32)
        self._dev = {}

    # ---- constructors -------------------------------------------------------------------------
    @staticmethod
    def self_pair_mask(parent=PART_PARENT):
        """eval_coll.py:616-620: pairs s < t that are neither identical nor parent / child."""
        n = len(parent)
        m = np.zeros((n, n), np.uint8)
        for s in range(n):
            for t in range(s + 1, n):
                if parent[s] != t and parent[t] != s:
                    m[s, t] = 1
        return m

    @classmethod
    def from_face_labels(cls, faces, face_part, n_verts, parent=PART_PARENT):
        """Parts given as a labelling of one closed mesh's faces: every boundary loop between two parts is
        capped with a fan to its centre (what MANO_PART.pkl stores explicitly for real MANO)."""
        faces = np.asarray(faces, np.int64)
        face_part = np.asarray(face_part)
        n_parts = int(face_part.max()) + 1
        caps, cap_id, part_faces = [], {}, []
        for p in range(n_parts):
            fp = faces[face_part == p]
            edges = {}
            for tri in fp:
                for a, b in ((tri[0], tri[1]), (tri[1], tri[2]), (tri[2], tri[0])):
                    edges.setdefault((min(a, b), max(a, b)), []).append((a, b))
            boundary = [e[0] for e in edges.values() if len(e) == 1]
            # group boundary edges into loops (connected components over shared vertices)
            parent_v = {}

            def find(x):
                while parent_v.setdefault(x, x) != x:
                    parent_v[x] = parent_v[parent_v[x]]
                    x = parent_v[x]
                return x

            for a, b in boundary:
                parent_v[find(a)] = find(b)
            loops = {}
            for a, b in boundary:
                loops.setdefault(find(a), []).append((a, b))
            extra = []
            for loop_edges in loops.values():
                vs = sorted({int(v) for e in loop_edges for v in e})
                key = tuple(vs)
                if key not in cap_id:
                    cap_id[key] = len(caps)
                    caps.append(vs)
                c = n_verts + cap_id[key]
                extra += [(b, a, c) for a, b in loop_edges]          # close the hole, consistent winding
            part_faces.append(np.concatenate([fp, np.asarray(extra, np.int64).reshape(-1, 3)]))
        return cls(n_verts, caps, part_faces, cls.self_pair_mask(parent))

    @classmethod
    def from_mano_part(cls, model_part, edge_vertex_id, n_verts=779, parent=PART_PARENT):
        """The reference's own tables: MANO_PART.pkl ('v-i' water-mesh vertex ids, 'f-i' local faces;
        eval_coll.py:139-148) and the cap loops hard-coded at :350-363."""
        n = len(parent)
        part_faces = [np.asarray(model_part["v-%d" % i])[np.asarray(model_part["f-%d" % i])] for i in range(n)]
        return cls(n_verts, edge_vertex_id, part_faces, cls.self_pair_mask(parent))

    @classmethod
    def synthetic_hand(cls):
        """Parts of the synthetic hand fixture (dsf_b200/assets/hand_topology.npz) incl. the wrist fan."""
        from .synthetic import make_synthetic_mano
        topo = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "hand_topology.npz"))
        model = make_synthetic_mano(0)
        f778 = np.asarray(model["f"], np.int64)
        ring = [121, 214, 215, 279, 239, 234, 92, 38, 122, 118, 117, 119, 120, 108, 79, 78]   # mano_layer.py:103-105
        fan = [(ring[i], ring[(i + 1) % 16], 778) for i in range(16)]
        faces = np.concatenate([f778, np.asarray(fan, np.int64)])
        label = np.concatenate([np.asarray(topo["face_part"]), np.zeros(16, np.int64)])
        return cls.from_face_labels(faces, label, 779)

    @classmethod
    def pair(cls, obj_faces, n_obj_verts, hand_faces):
        """intersect_vox(obj_mesh, hand_mesh): vertices = [object | hand]; voxels of the object inside the hand."""
        hand_faces = np.asarray(hand_faces, np.int64) + n_obj_verts
        n = int(max(np.max(obj_faces) + 1, n_obj_verts, hand_faces.max() + 1))
        return cls(n, [], [obj_faces, hand_faces], [[0, 0], [1, 0]])

    # ---- device copies --------------------------------------------------------------------------
    def device(self, dev):
        key = str(dev)
        if key not in self._dev:
            t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)
            self._dev[key] = dict(cap_ptr=t(self.cap_ptr, torch.int32), cap_idx=t(self.cap_idx, torch.int32),
                                  part_ptr=t(self.part_ptr, torch.int32), faces=t(self.faces, torch.int32),
                                  mask=t(self.pair_mask, torch.uint8))
        return self._dev[key]

    def is_watertight(self):
        for f in self.part_faces:
            e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), 1)
            _, cnt = np.unique(e, axis=0, return_counts=True)
            if not (cnt == 2).all():
                return False
        return True


def intersect_counts(verts, topo: PartTopology, pitch=2.0):
    """verts (B,n_verts,3) CUDA fp32 (mm) -> dict(volume (B,) f64, pair_counts (B,P,P) i64,
    voxel_counts (B,P) i64, status (B,) i32).  No host synchronisation."""
    lib = L.lib()
    v = L.f32c(verts)
    if v.dim() != 3 or v.shape[1] != topo.n_verts or v.shape[2] != 3:
        raise ValueError(f"verts must be (B,{topo.n_verts},3)")
    B, dev, P = v.shape[0], v.device, topo.n_parts
    d = topo.device(dev)
    pc = torch.empty(B, P, P, dtype=torch.int64, device=dev)
    vc = torch.empty(B, P, dtype=torch.int64, device=dev)
    vol = torch.empty(B, dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    ws = torch.empty(lib.dsf_intersect_workspace_bytes(B, topo.n_verts, len(topo.cap_loops), P) // 8 + 1,
                     dtype=torch.float64, device=dev)
    L.check(lib.dsf_intersect_vox(B, topo.n_verts, v.data_ptr(), len(topo.cap_loops), d["cap_ptr"].data_ptr(),
                                  d["cap_idx"].data_ptr(), P, d["part_ptr"].data_ptr(), d["faces"].data_ptr(),
                                  d["mask"].data_ptr(), float(pitch), pc.data_ptr(), vc.data_ptr(), vol.data_ptr(),
                                  status.data_ptr(), ws.data_ptr(), L.stream_ptr()))
    return dict(volume=vol, pair_counts=pc, voxel_counts=vc, status=status)


def self_intersection(meshes, topo: PartTopology | None = None, pitch=2):
    """``eval_coll.self_intersection(part_mesh_list, pitch)`` for a batch: meshes (B,779,3) in mm ->
    (B,) intersection volumes in mm^3 (float64).  Raises like trimesh does when a face needs more than
    10 subdivision levels."""
    topo = topo or PartTopology.synthetic_hand()
    out = intersect_counts(meshes, topo, pitch)
    if bool((out["status"] != 0).any()):
        raise ValueError("voxelisation failed (max_iter exceeded or part larger than the voxel bitmap)")
    return out["volume"]


def intersect_vox(obj_verts, obj_faces, hand_verts, hand_faces, pitch=2):
    """``util/intersect.py:intersect_vox(obj_mesh, hand_mesh, pitch)`` for a batch of scenes:
    obj_verts (B,No,3), hand_verts (B,Nh,3) -> (B,) volumes."""
    topo = PartTopology.pair(np.asarray(obj_faces), obj_verts.shape[1], np.asarray(hand_faces))
    verts = torch.cat([obj_verts, hand_verts], 1)
    out = intersect_counts(verts, topo, pitch)
    if bool((out["status"] != 0).any()):
        raise ValueError("voxelisation failed (max_iter exceeded or object larger than the voxel bitmap)")
    return out["volume"]
