#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own code (imported unmodified from
/root/reference via oracle/ref_import.py) on the synthetic MANO model.  Run in the build
container only; the outputs are committed so the GPU box never needs /root/reference.

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from dsf_b200.synthetic import write_mano_pkl, sample_fit_inputs  # noqa: E402
from oracle.ref_import import import_reference_mano_module, make_reference_render  # noqa: E402

NYU = (588.03, 587.07, 320.0, 240.0)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    mod = import_reference_mano_module()
    tmp = tempfile.mkdtemp()
    pkl = write_mano_pkl(tmp, seed=0)
    ref = mod.MANO_SMPL(pkl, "nyu")

    B = 6
    inp = sample_fit_inputs(B, seed=1)
    p = torch.tensor(inp["params"])
    p[0, :48] = 0.0                      # zero-pose known answer (SURVEY section 4 invariant 1)
    p[1, 3:48] = 0.0
    p.requires_grad_(True)
    quat, theta, beta, cam = p[:, :3], p[:, 3:48], p[:, 48:58], p[:, 58:]

    verts_m, joints_m, Rs = ref.forward(beta, theta, quat, get_skin=True)
    verts, joints = ref.get_mano_vertices(quat, theta, beta, cam, global_scale=1 / 125)
    # a fixed random cotangent gives one gradient vector per parameter
    gv = torch.randn(verts.shape, generator=torch.Generator().manual_seed(3))
    gj = torch.randn(joints.shape, generator=torch.Generator().manual_seed(4))
    (g_p,) = torch.autograd.grad((verts * gv).sum() + (joints * gj).sum(), p, retain_graph=True)

    coll = ref.calculate_coll(joints, verts.detach())
    # make penetration likely for a second case: shrink the hand so spheres overlap
    (g_coll,) = torch.autograd.grad(coll, p, retain_graph=True, allow_unused=True)
    sph_c, sph_r = ref.get_sphere_radius(joints.clone(), verts.detach())
    # per-hand collision totals before the gate / batch mean
    cen, rad = sph_c, sph_r
    d = cen[:, :, None] - cen[:, None]
    dis = torch.sqrt((d * d).sum(-1) + 1e-8)
    err = torch.clamp(rad[:, :, None] + rad[:, None] - dis, min=0) * ref.mask
    coll_per_hand = err.sum((-1, -2))

    # 4-component quaternion branch (mano_layer.py:602-609)
    q4 = torch.randn(B, 4, generator=torch.Generator().manual_seed(5))
    v4, j4, _ = ref.forward(beta.detach(), theta.detach(), q4, get_skin=True)

    # Render's pure-torch helpers
    rnd = make_reference_render(mod, ref, NYU, (640, 480), (128, 128))
    center3d = torch.tensor(inp["center3d"])
    cube = torch.tensor(inp["cube"])
    center3d[0] = torch.tensor([0.0, 0.0, 800.0])          # SURVEY section 4 invariant 7
    cube[0] = 250.0
    center2d = rnd.points3DToImg(center3d.unsqueeze(1)).squeeze(1)
    xs, xe, ys, ye, zs, ze = rnd.comToBounds(center2d, cube)
    M = rnd.Offset2Trans(xs, xe, ys, ye)
    hand_j = joints.detach() * cube.unsqueeze(1) / 2 + center3d.unsqueeze(1)
    joint_uvd = rnd.JointTrans(hand_j, M, center2d, cube)

    # literal pixel chain: push an index image through resize + warpPerspective
    S, W, H = 640, 640, 480
    idx_img = (torch.arange(S * S, dtype=torch.float32) + 1).view(1, 1, S, S).repeat(B, 1, 1, 1)
    resized = rnd.resize(idx_img)
    cropped = rnd.warpPerspective(resized, M)
    src = cropped.round().long().view(B, 128, 128) - 1          # -1 = zero padding
    # normalize_img on a small synthetic depth patch
    zimg = torch.tensor([[0.0, -1.0, 700.0, 790.0, 800.0, 930.0, 1500.0, 100.0]]).view(1, 1, 1, 8).repeat(B, 1, 1, 1)
    zimg = zimg + (center3d[:, 2] - 800.0).view(B, 1, 1, 1) * (zimg > 0)
    znorm = rnd.normalize_img(zimg.clone(), center2d, cube)

    # crop_hand of the data loader (data/render_loader.py:1209-1227), run on the reference's own class
    from oracle.ref_import import import_reference_loader_module
    from oracle import raster_oracle as ro
    from oracle import mano_oracle as mo
    rl = import_reference_loader_module()
    ld = rl.loader.__new__(rl.loader)
    ld.img_size, ld.paras, ld.flip = 128, NYU, 1
    consts = mo.ManoConstants(__import__("dsf_b200").make_synthetic_mano(0))
    vw = verts.detach() * cube[:, None] / 2 + center3d[:, None]
    view, xs_, ys_, M_d = ro.make_view("direct", center3d, cube, NYU, 640, 480, 128)
    _, zb, _, _ = ro.render(vw, consts.faces, view, xs_, ys_)
    crop_in = ro.normalize_depth(zb, view)[:, None]
    teacher = joints.detach() + 0.05 * torch.randn(joints.shape, generator=torch.Generator().manual_seed(9))
    teacher[:, :, 2] *= 0.6                                  # a tighter z box so the crop actually bites
    crop_out = ld.crop_hand(crop_in.clone(), teacher, center3d, M_d, cube)

    # seg_pcl (mano_layer.py:404-426) on the reference class
    seg_pts = verts.detach()[:, torch.randint(0, 778, (300,), generator=torch.Generator().manual_seed(11))] \
        + 0.08 * torch.randn(B, 300, 3, generator=torch.Generator().manual_seed(12))
    seg_joints = joints.detach() + 0.02 * torch.randn(joints.shape, generator=torch.Generator().manual_seed(13))
    seg_ref = ref.seg_pcl(seg_joints, joints.detach(), verts.detach(), seg_pts)

    out = os.path.join(ROOT, "tests", "golden", "mano_golden.npz")
    np.savez_compressed(
        out,
        params=p.detach().numpy(), verts_m=verts_m.detach().numpy(), joints_m=joints_m.detach().numpy(),
        Rs=Rs.detach().numpy(), verts=verts.detach().numpy(), joints=joints.detach().numpy(),
        gv=gv.numpy(), gj=gj.numpy(), g_params=g_p.numpy(),
        coll=coll.detach().numpy(), g_coll=g_coll.numpy(), coll_per_hand=coll_per_hand.detach().numpy(),
        sph_c=sph_c.detach().numpy(), sph_r=sph_r.detach().numpy(), coll_mask=ref.mask.numpy(),
        q4=q4.numpy(), verts_q4=v4.detach().numpy(), joints_q4=j4.detach().numpy(),
        faces=ref.faces.numpy().astype(np.int32),
        joint_faces_len=np.array([len(f) for f in ref.joint_faces]),
        center3d=center3d.numpy(), cube=cube.numpy(), center2d=center2d.numpy(),
        bounds=torch.stack([xs, xe, ys, ye], 1).numpy(), M=M.numpy(), joint_uvd=joint_uvd.detach().numpy(),
        literal_src=src.numpy().astype(np.int32), zimg=zimg.numpy(), znorm=znorm.numpy(),
        crop_in=crop_in.numpy(), crop_teacher=teacher.numpy(), crop_M=M_d.numpy(), crop_out=crop_out.numpy(),
        seg_pts=seg_pts.numpy(), seg_joints=seg_joints.numpy(), seg_ref=seg_ref.numpy().astype(np.int32),
    )
    print("crop_hand removed px:", int(((crop_in < 0.99) & (crop_out >= 0.99)).sum()), "of", int((crop_in < 0.99).sum()))
    print("wrote", out, os.path.getsize(out), "bytes")
    print("coll", float(coll), "per hand", coll_per_hand.detach().numpy())
    print("mask ones", int(ref.mask.sum()), "sym", bool((ref.mask == ref.mask.T).all()))
    print("M[0]", M[0].numpy(), "bounds0", xs[0].item(), xe[0].item(), ys[0].item(), ye[0].item())


if __name__ == "__main__":
    main()
