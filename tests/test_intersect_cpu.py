"""Row I1 (intersection volume): the CPU restatement against analytic cases and against an independent
numpy implementation (direct lattice voxelisation + generalised winding number).  Parity with trimesh
itself is unpinned (not installed, no reference vectors) - see oracle/intersect_oracle.c."""
import numpy as np
import torch

from dsf_b200.intersection import PART_PARENT, PartTopology
from oracle import intersect_oracle as io

from oracle.shapes import CUBE_F, cube, icosphere  # noqa: E402


def two_part_topology(fa, na, fb):
    """part 0 = A, part 1 = B; count voxels of B (t = 1) inside A (s = 0)."""
    return PartTopology(na + int(fb.max()) + 1, [], [fa, fb + na], [[0, 1], [0, 0]])


def test_offset_cubes_analytic():
    va, vb = cube(0.3, 10.3), cube(5.3, 15.3)
    topo = two_part_topology(CUBE_F, 8, CUBE_F)
    vol, pc, vc = io.intersect_vox(np.concatenate([va, vb])[None], topo, 1.0)
    assert vc[0].tolist() == [11 ** 3 - 9 ** 3, 11 ** 3 - 9 ** 3]      # index-cube surfaces
    assert pc[0, 0, 1] == 6 ** 3 - 5 ** 3 and vol[0] == 91.0
    vol2, pc2, vc2 = io.intersect_vox(np.concatenate([va, vb])[None], topo, 2.0)
    assert vol2[0] == pc2[0, 0, 1] * 8.0 and 0 < pc2[0, 0, 1] < pc[0, 0, 1]


def test_nested_and_disjoint_spheres():
    vs, fs = icosphere(20.0, (3.1, -2.2, 400.4), 3)
    vi, fi = icosphere(6.0, (5.0, 1.0, 398.0), 2)
    vd, _ = icosphere(6.0, (60.0, 1.0, 398.0), 2)
    topo = two_part_topology(fs, len(vs), fi)
    vol, pc, vc = io.intersect_vox(np.stack([np.concatenate([vs, vi]), np.concatenate([vs, vd])]), topo, 2.0)
    assert pc[0, 0, 1] == vc[0, 1] > 50            # every surface voxel of the inner sphere is inside
    assert pc[1, 0, 1] == 0 and vol[1] == 0.0
    # surface voxel count scales like area / pitch^2
    assert 0.5 < vc[0, 0] / (4 * np.pi * 20.0 ** 2 / 4.0) < 2.0


def _lattice_voxels(wv, faces, pitch):
    keys = set()
    for a, b, c in wv[faces]:
        e = max(np.linalg.norm(b - a), np.linalg.norm(c - b), np.linalg.norm(a - c))
        k = 0
        while e > pitch / 2:
            e /= 2
            k += 1
        n = 1 << k
        ii, jj = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
        m = ii + jj <= n
        ii, jj = ii[m], jj[m]
        p = (a[None] * (n - ii - jj)[:, None] + b[None] * ii[:, None] + c[None] * jj[:, None]) / n
        keys.update(map(tuple, np.rint(p / pitch).astype(np.int64)))
    return np.array(sorted(keys), np.float64) * pitch


def _winding_inside(p, tri):
    a, b, c = tri[:, 0][None] - p[:, None], tri[:, 1][None] - p[:, None], tri[:, 2][None] - p[:, None]
    la, lb, lc = (np.linalg.norm(x, axis=-1) for x in (a, b, c))
    num = np.einsum("ijk,ijk->ij", a, np.cross(b, c))
    den = la * lb * lc + (a * b).sum(-1) * lc + (b * c).sum(-1) * la + (c * a).sum(-1) * lb
    w = (2 * np.arctan2(num, den)).sum(1) / (4 * np.pi)
    # crossing parity (what trimesh's contains decides) = parity of the winding number; the strongly
    # curled test hands have parts that overlap themselves (winding 2) or turn inside out (-1)
    return np.rint(w).astype(np.int64) % 2 == 1


def test_hand_parts_against_independent_numpy_implementation(mano_model):
    """synthetic hand with curled fingers: per-pair counts from the C restatement (recursive midpoint
    subdivision + ray parity) == direct lattice voxelisation + winding number in numpy."""
    from dsf_b200 import sample_fit_inputs
    from oracle import mano_oracle as mo

    topo = PartTopology.synthetic_hand()
    assert topo.n_parts == 15 and len(topo.cap_loops) == 14 and topo.is_watertight()
    assert sorted(len(l) for l in topo.cap_loops) == sorted([10, 10, 10, 10, 10, 10, 10, 10, 10, 9, 10, 10, 11, 10])
    assert int(topo.pair_mask.sum()) == 91        # 105 pairs minus the 14 parent / child ones (eval_coll.py:615)
    c = mo.ManoConstants(mano_model)
    p = torch.from_numpy(sample_fit_inputs(3, seed=3)["params"])
    quat, theta, beta, cam = mo.split_params(p)
    v, _ = mo.get_mano_vertices(c, quat, theta * 3, beta, cam)
    v = v.detach().numpy()
    vol, pc, vc = io.intersect_vox(v, topo, 2.0)
    assert (vol > 0).all() and (vol == pc.sum((1, 2)) * 8.0).all()
    b = 0
    wv = np.concatenate([v[b].astype(np.float64)] + [v[b][l].astype(np.float64).mean(0, keepdims=True) for l in topo.cap_loops])
    pts = [_lattice_voxels(wv, f, 2.0) for f in topo.part_faces]
    assert [len(q) for q in pts] == vc[b].tolist()
    mism = 0
    for s in range(15):
        for t in range(15):
            if not topo.pair_mask[s, t]:
                assert pc[b, s, t] == 0
                continue
            f = topo.part_faces[s]
            lo, hi = wv[f].reshape(-1, 3).min(0), wv[f].reshape(-1, 3).max(0)
            q = pts[t][((pts[t] >= lo) & (pts[t] <= hi)).all(1)]
            n = int(_winding_inside(q, wv[f]).sum()) if len(q) else 0
            mism += abs(n - int(pc[b, s, t]))
    assert mism == 0, mism
