"""Row I1 on the GPU: dsf_intersect_vox against the CPU restatement (integer counts: exact)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _curled_hands(mano_model, B, seed, curl):
    from dsf_b200 import sample_fit_inputs
    from oracle import mano_oracle as mo

    c = mo.ManoConstants(mano_model)
    p = torch.from_numpy(sample_fit_inputs(B, seed=seed)["params"])
    quat, theta, beta, cam = mo.split_params(p)
    v, _ = mo.get_mano_vertices(c, quat, theta * curl, beta, cam)
    return v.detach().contiguous()


@pytest.mark.parametrize("pitch,B", [(2.0, 48), (1.0, 6)])
def test_self_intersection_matches_oracle(mano_model, pitch, B):
    from dsf_b200.intersection import PartTopology, intersect_counts, self_intersection
    from oracle import intersect_oracle as io

    topo = PartTopology.synthetic_hand()
    v = torch.cat([_curled_hands(mano_model, B // 2, 5, 3.0), _curled_hands(mano_model, B - B // 2, 6, 1.0)])
    ref_vol, ref_pc, ref_vc = io.intersect_vox(v.numpy(), topo, pitch)
    out = intersect_counts(v.cuda(), topo, pitch)
    assert (out["status"] == 0).all()
    assert np.array_equal(out["voxel_counts"].cpu().numpy(), ref_vc)
    assert np.array_equal(out["pair_counts"].cpu().numpy(), ref_pc)
    assert np.array_equal(out["volume"].cpu().numpy(), ref_vol)
    assert (ref_vol > 0).sum() >= B // 2 and (ref_pc > 0).sum() > B          # the case is not vacuous
    vol = self_intersection(v.cuda(), topo, pitch)
    assert torch.equal(vol, out["volume"])
    # a rigid translation by a whole number of voxels moves the voxel grid with the mesh
    shift = torch.tensor([4 * pitch, -6 * pitch, 10 * pitch])
    out2 = intersect_counts((v + shift).cuda(), topo, pitch)
    same = (torch.as_tensor(ref_pc).cuda() == out2["pair_counts"]).float().mean()
    assert same > 0.97        # float32 vertices re-round after the shift: a few boundary voxels may move


def test_intersect_vox_object_hand_pair_and_analytic_cubes():
    from dsf_b200.intersection import intersect_vox
    from oracle.shapes import CUBE_F, cube, icosphere

    va, vb = torch.tensor(cube(0.3, 10.3))[None], torch.tensor(cube(5.3, 15.3))[None]
    vol = intersect_vox(vb.cuda(), CUBE_F, va.cuda(), CUBE_F, pitch=1)        # voxels of B inside A
    assert vol.item() == 91.0
    vs, fs = icosphere(20.0, (3.1, -2.2, 400.4), 3)
    vi, fi = icosphere(6.0, (5.0, 1.0, 398.0), 2)
    vd, _ = icosphere(6.0, (60.0, 1.0, 398.0), 2)
    obj = torch.tensor(np.stack([vi, vd])).cuda()
    hand = torch.tensor(np.stack([vs, vs])).cuda()
    vol = intersect_vox(obj, fi, hand, fs, pitch=2).cpu()
    assert vol[0] > 50 * 8 and vol[1] == 0


def test_intersect_status_flags():
    from dsf_b200.intersection import PartTopology, intersect_counts, self_intersection
    from oracle.shapes import CUBE_F, cube

    topo = PartTopology(16, [], [CUBE_F, CUBE_F + 8], [[0, 1], [0, 0]])
    big = torch.tensor(np.concatenate([cube(0, 10), cube(0, 4000)]))[None].cuda()
    out = intersect_counts(big, topo, 1.0)
    assert int(out["status"][0]) & 1 and float(out["volume"][0]) == -1.0      # one z-layer larger than the bitmap
    with pytest.raises(ValueError):
        self_intersection(big, topo, 1.0)
    tiny = torch.tensor(np.concatenate([cube(0, 10), cube(0, 8)]))[None].cuda()
    out = intersect_counts(tiny, topo, 0.02)                                   # > 10 subdivision levels
    assert int(out["status"][0]) & 2
    with pytest.raises(ValueError):
        intersect_counts(big[:, :5], topo, 1.0)
