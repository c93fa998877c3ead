"""GPU tests of the rows the round-1 review listed as missing or only shape-tested: the synthetic-data generator
extras (mask_img, synth2real, surface_loss / chamfer), the rarely used sphere-model variants, the resampling
helpers, the pytorch3d-shaped Fragments adapter and every Render entry point against the oracle.  Golden vectors
come from the reference's own code (tests/golden/make_golden_extras.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NYU = (588.03, 587.07, 320.0, 240.0)


@pytest.fixture(scope="module")
def ex():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "extras_golden.npz")))


@pytest.fixture(scope="module")
def rnd(mano_model):
    from dsf_b200.mano_layer import Render

    torch.cuda.set_device(0)
    return Render(mano_model, "nyu", NYU, (640, 480), (128, 128))


@pytest.fixture(scope="module")
def crop_in(golden):
    return torch.tensor(golden["crop_in"]).cuda()


def _inputs(B, seed):
    from dsf_b200 import sample_fit_inputs

    return {k: torch.from_numpy(v) for k, v in sample_fit_inputs(B, seed=seed).items()}


# ------------------------------------------------------------------------------------------------
# f4: mask_img / synth2real / surface_loss
# ------------------------------------------------------------------------------------------------
def test_mask_img_matches_reference_golden(rnd, ex, crop_in):
    # same seeds, same generator calls as the reference -> the same occluders, pixel for pixel
    np.random.seed(11)
    torch.manual_seed(11)
    out = rnd.mask_img(crop_in, torch.tensor(ex["mask_joint_uvd"]).cuda(), 0.15, 0.3)
    ref = torch.tensor(ex["mask_out"])
    assert out.shape == ref.shape
    assert (ref != torch.tensor(crop_in.cpu())).float().mean() > 0.005, "the occluders must hit the hand"
    assert torch.equal(out.cpu(), ref)
    # the deterministic kernel on the spheres the reference drew
    out2 = rnd.mask_spheres(crop_in, torch.tensor(ex["mask_centres"]).cuda(), torch.tensor(ex["mask_radii"]).cuda())
    assert torch.equal(out2.cpu(), ref)
    # no spheres: identity; a huge sphere: everything background
    B = crop_in.shape[0]
    assert torch.equal(rnd.mask_spheres(crop_in, torch.zeros(B, 0, 3).cuda(), torch.zeros(B, 0).cuda()), crop_in)
    allbg = rnd.mask_spheres(crop_in, torch.zeros(B, 1, 3).cuda(), torch.full((B, 1), 10.0).cuda())
    assert (allbg == 1.0).all()


def test_synth2real_matches_reference_golden(rnd, ex, crop_in):
    from dsf_b200 import _lib as L

    lib = L.lib()
    B, R = crop_in.shape[0], crop_in.shape[-1]
    # the Gaussian alone (no noise) against GaussianSmoothing(5) on the reflect-padded image
    for sigma in (1.7, 0.5):
        out = torch.empty_like(crop_in)
        L.check(lib.dsf_synth2real(B, R, crop_in.data_ptr(), None, 1, 0.95, sigma, out.data_ptr(), L.stream_ptr()))
        np.testing.assert_allclose(out.cpu().numpy(), ex[f"smooth_{sigma}"], rtol=0, atol=2e-6)
    # the whole function with the reference's random stream
    torch.manual_seed(3)
    np.testing.assert_allclose(rnd.synth2real(crop_in).cpu().numpy(), ex["s2r_default"], rtol=0, atol=3e-6)
    torch.manual_seed(4)
    np.testing.assert_allclose(rnd.synth2real(crop_in, noise=0.02, noise_patch=4, sigma=0.5).cpu().numpy(),
                               ex["s2r_p4_s05"], rtol=0, atol=3e-6)
    torch.manual_seed(5)
    got = rnd.synth2real(crop_in, noise=0.05, noise_patch=2, sigma=0).cpu()
    assert torch.equal(got, torch.tensor(ex["s2r_nosmooth"]))
    assert torch.equal(got[crop_in.cpu() >= 0.95], crop_in.cpu()[crop_in.cpu() >= 0.95])     # background untouched


def test_chamfer_and_surface_loss(rnd, crop_in, golden):
    from dsf_b200.render_loss import chamfer_distance, depth_loss, surface_loss  # noqa: F401  (train_render.py:16)

    gen = torch.Generator().manual_seed(0)
    for B, P1, P2 in ((3, 1024, 779), (2, 37, 1500), (1, 1, 1)):
        x = torch.randn(B, P1, 3, generator=gen).cuda().requires_grad_(True)
        y = torch.randn(B, P2, 3, generator=gen).cuda().requires_grad_(True)
        loss, _ = chamfer_distance(x, y)
        gx, gy = torch.autograd.grad(loss, (x, y))
        xd, yd = x.detach().double().requires_grad_(True), y.detach().double().requires_grad_(True)
        d = ((xd[:, :, None] - yd[:, None]) ** 2).sum(-1)                       # published definition, float64
        ref = d.min(2)[0].mean(1).mean() + d.min(1)[0].mean(1).mean()
        gxr, gyr = torch.autograd.grad(ref, (xd, yd))
        assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
        assert (gx.double() - gxr).abs().max() <= 1e-4 * gxr.abs().max() + 1e-12
        assert (gy.double() - gyr).abs().max() <= 1e-4 * gyr.abs().max() + 1e-12
    # surface_loss on real crops: runs, is non-negative, is zero-gradient-free of NaNs, falls back to verts on an empty crop
    sl = surface_loss()
    center, cube, M = (torch.tensor(golden[k]).cuda() for k in ("center3d", "cube", "crop_M"))
    B = crop_in.shape[0]
    verts = (torch.randn(B, 779, 3, generator=gen) * 0.3).cuda().requires_grad_(True)
    loss = sl(crop_in, None, verts, None, center, M, cube)
    (g,) = torch.autograd.grad(loss, verts)
    assert loss.item() > 0 and torch.isfinite(g).all() and g.abs().max() > 0
    empty = crop_in.clone()
    empty[1] = 1.0
    assert sl.Img2pcl(empty, center, M, cube, verts) is verts
    assert sl(empty, None, verts, None, center, M, cube).item() == 0.0


# ------------------------------------------------------------------------------------------------
# resampling helpers and sphere-model variants vs the reference golden
# ------------------------------------------------------------------------------------------------
def test_resampling_helpers_match_reference_golden(rnd, ex, crop_in):
    S = 640
    idx_img = (torch.arange(S * S, dtype=torch.float32).view(1, 1, S, S).repeat(2, 1, 1, 1) % 4099.0).cuda()
    sensor = rnd.resize(idx_img)
    assert sensor.shape == (2, 1, 480, 640)
    got, ref = sensor[:, :, ::7, ::5].cpu().numpy(), ex["resize_sub"]
    # sensor row r reads raster row (2 r + 1) * 2 / 3 - 1 / 2: an exact .5 for r = 1 (mod 3), which grid_sample
    # rounds by float noise (tie class T3; CPU and GPU builds of torch disagree there) - every other row is exact
    rows = np.arange(0, 480, 7)
    clean = rows % 3 != 1
    assert np.array_equal(got[:, :, clean], ref[:, :, clean])
    assert (got != ref).mean() < 0.34
    M = torch.tensor(ex["helpers_M"]).cuda()
    np.testing.assert_allclose(rnd.affine_grid(sensor, M)[:, ::9, ::9].cpu().numpy(), ex["affine_grid_sub"], atol=2e-5)
    ref_sensor = sensor.clone()
    warp = rnd.warpPerspective(ref_sensor, M).cpu().numpy()
    assert warp.shape == ex["warp"].shape and (warp != ex["warp"]).mean() < 0.05
    pos = torch.where(crop_in < 0.99, crop_in + 2.0, torch.zeros_like(crop_in))
    np.testing.assert_allclose(rnd.massCenter(pos).cpu().numpy(), ex["mass_center"], rtol=1e-5)
    out = rnd.ResizeRenderImg(idx_img[:, :, :64, :64].repeat(1, 1, 10, 10))
    assert out.shape == (2, 1, 480, 640) and torch.isfinite(out).all()


def test_sphere_model_variants_match_reference_golden(rnd, ex):
    layer = rnd.mano_layer
    p = torch.tensor(ex["sv_params"]).cuda()
    verts, joints = layer.get_mano_vertices(p[:, :3], p[:, 3:48], p[:, 48:58], p[:, 58:], global_scale=1 / 125)
    verts, joints = verts.detach(), joints.detach()
    pcl, j_pwe = torch.tensor(ex["sv_pcl"]).cuda(), torch.tensor(ex["sv_joints_pwe"]).cuda()
    np.testing.assert_allclose(layer.get_sphere(joints).cpu().numpy(), ex["sv_get_sphere"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(layer.get_radius(joints, verts).cpu().numpy(), ex["sv_get_radius"], rtol=1e-4, atol=1e-6)
    for key, jc in (("sv_pwe_coll", j_pwe), ("sv_pwe_coll_tight", joints * 0.6)):
        got = layer.calculate_PWE_coll(jc, joints, verts).item()
        assert abs(got - float(ex[key])) <= 1e-4 * abs(float(ex[key])) + 1e-7, (key, got, float(ex[key]))
    assert float(ex["sv_pwe_coll_tight"]) > 0, "the tight configuration must make spheres overlap"
    # gradient reaches the PWE joints (centres) through the torch interpolation
    jg = (joints * 0.6).clone().requires_grad_(True)
    (g,) = torch.autograd.grad(layer.calculate_PWE_coll(jg, joints, verts), jg)
    assert torch.isfinite(g).all() and g.abs().max() > 0
    # labels: equal except where two sphere surfaces are equidistant to within rounding
    for key, fn in (("sv_seg15", layer.seg_pcl), ("sv_seg21", layer.seg_pcl_21), ("sv_seg5", layer.seg_pcl_finger)):
        got = fn(j_pwe, joints, verts, pcl).cpu().numpy()
        assert (got != ex[key]).mean() < 5e-3, key
    seg15, seg21, seg5 = (torch.tensor(ex[k]).cuda() for k in ("sv_seg15", "sv_seg21", "sv_seg5"))
    d15, idx15 = layer.calculate_point2shpere_distance(joints, verts, pcl, seg15)
    np.testing.assert_allclose(d15.cpu().numpy(), ex["sv_p2s15"], rtol=2e-4, atol=1e-6)
    assert (idx15.cpu().numpy() != ex["sv_p2s15_idx"]).mean() < 5e-3
    np.testing.assert_allclose(layer.calculate_point2shpere_distance_21(joints, verts, pcl, seg21).cpu().numpy(),
                               ex["sv_p2s21"], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(layer.calculate_point2shpere_distance_finger(joints, verts, pcl, seg5).cpu().numpy(),
                               ex["sv_p2s5"], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(layer.calculate_point2mesh_distance(verts, pcl, seg15).cpu().numpy(), ex["sv_p2m"],
                               rtol=2e-4, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# pytorch3d-shaped boundary
# ------------------------------------------------------------------------------------------------
def test_fragments_adapter(rnd, mano_model):
    from dsf_b200.fragments import Fragments, MeshRasterizer
    from oracle import mano_oracle as mo
    from oracle import raster_oracle as ro

    c32 = mo.ManoConstants(mano_model)
    B = 3
    inp = _inputs(B, seed=15)
    p, c3, cube = inp["params"].cuda(), inp["center3d"].cuda(), inp["cube"].cuda()
    joints, mesh = rnd.get_mesh_xyz(p)
    vw = (mesh * cube[:, None] / 2 + c3[:, None]).contiguous()
    fr = MeshRasterizer(rnd)(vw, c3, cube)
    assert isinstance(fr, Fragments)
    assert fr.pix_to_face.shape == (B, 128, 128, 1) and fr.pix_to_face.dtype == torch.int64
    assert fr.zbuf.shape == (B, 128, 128, 1) and fr.bary_coords.shape == (B, 128, 128, 1, 3) and fr.dists.shape == (B, 128, 128, 1)
    view, xs, ys, _ = rnd._view(c3, cube)
    p_ref, z_ref, b_ref, _ = ro.render(vw.cpu(), c32.faces, view[:, :8].cpu().contiguous(), xs.cpu(), ys.cpu(), want_bary=True)
    F = c32.faces.shape[0]
    packed_ref = torch.where(p_ref >= 0, p_ref.long() + torch.arange(B).view(B, 1, 1) * F, p_ref.long())
    assert torch.equal(fr.pix_to_face[..., 0].cpu(), packed_ref)             # pytorch3d's packed face index
    assert torch.equal(fr.zbuf[..., 0].cpu(), z_ref) and torch.equal(fr.bary_coords[:, :, :, 0].cpu(), b_ref)
    bg = fr.pix_to_face[..., 0] < 0
    assert (fr.zbuf[..., 0][bg] == -1).all() and (fr.dists[..., 0][bg] == -1).all() and (fr.dists[..., 0][~bg] <= 0).all()


# ------------------------------------------------------------------------------------------------
# R5: every Render entry point against the oracle (literal pixel chain)
# ------------------------------------------------------------------------------------------------
def _oracle_literal(c32, verts_cam, c3, cube, M=None, R=128):
    """normalised depth crop of the oracle for camera-space vertices; M given = M_render / getDepth's own crop"""
    from oracle import mano_oracle as mo
    from oracle import raster_oracle as ro

    view, xs, ys, M0 = ro.make_view("literal", c3, cube, NYU, 640, 480, R)
    if M is not None:
        xi, yi, tx, ty = mo.literal_sample_maps(M, 640, 480, 640, R)
        ndc = torch.cat([ro.pix_to_ndc(640), torch.tensor([float("nan")])])
        xs, ys = ndc[xi].contiguous(), ndc[yi].contiguous()
    else:
        _, _, tx, ty = mo.literal_sample_maps(M0, 640, 480, 640, R)
    p2f, z, _, _ = ro.render(verts_cam, c32.faces, view, xs, ys)
    return ro.normalize_depth(z, view), ty[:, :, None] | tx[:, None, :]


def _assert_same_image(img_gpu, img_ref, amb, what):
    d = (img_gpu[:, 0].cpu() - img_ref).abs().masked_fill(amb, 0.0)       # T3 rows / columns excluded
    assert (img_ref < 0.99).float().mean() > 0.02, what
    # vertices differ by ~1e-7 between the two MANO implementations: a handful of silhouette pixels may flip
    assert (d > 1e-5).float().mean() < 2e-3, (what, (d > 1e-5).float().mean())


def test_render_entry_points_match_oracle(rnd, mano_model):
    from oracle import mano_oracle as mo

    c32 = mo.ManoConstants(mano_model)
    B = 5
    inp = _inputs(B, seed=70)
    p, c3, cube = inp["params"], inp["center3d"], inp["cube"]
    pg, c3g, cubeg = p.cuda(), c3.cuda(), cube.cuda()
    q, t, b, cam = mo.split_params(p)
    v_n, j_n = mo.get_mano_vertices(c32, q, t, b, cam, global_scale=1 / 125)
    vw, jw = v_n * cube[:, None] / 2 + c3[:, None], j_n * cube[:, None] / 2 + c3[:, None]
    c2 = mo.points3d_to_img(c3, NYU)

    # mesh2img: the crop computed from (center3d, cube)
    ref, amb = _oracle_literal(c32, vw, c3, cube)
    _assert_same_image(rnd.mesh2img(vw.cuda(), c3g, cubeg), ref, amb, "mesh2img")
    from oracle import raster_oracle as ro
    _, _, _, M0 = ro.make_view("literal", c3, cube, NYU, 640, 480, 128)
    # normal_render places the hand with the older [0,1] convention, (x + 1) / 2 * cube + centre (:1049-1050)
    vw_n, jw_n = (v_n + 1) / 2 * cube[:, None] + c3[:, None], (j_n + 1) / 2 * cube[:, None] + c3[:, None]
    ref_n, amb_n = _oracle_literal(c32, vw_n, c3, cube)
    img_n, juvd_n, jxyz_n, vxyz_n = rnd.normal_render(pg, c3g, cubeg)
    assert img_n.shape == (B, 1, 128, 128) and torch.isfinite(img_n).all()
    d_n = (img_n[:, 0].cpu() - ref_n).abs().masked_fill(amb_n, 0.0)
    assert (d_n > 1e-5).float().mean() < 2e-3
    np.testing.assert_allclose(juvd_n.cpu().numpy(), mo.joint_trans(jw_n, M0, c2, cube, NYU, 128).numpy(), rtol=0, atol=3e-4)
    np.testing.assert_allclose(jxyz_n.cpu().numpy(), (j_n + 1).numpy(), rtol=0, atol=3e-5)
    np.testing.assert_allclose(vxyz_n.cpu().numpy(), (v_n + 1).numpy(), rtol=0, atol=3e-5)

    # getDepth with a caller-supplied crop transform (a shifted, tighter crop) and an extra view rotation
    M = M0.clone()
    M[:, 0, 0] *= 1.25
    M[:, 1, 1] *= 1.25
    M[:, 0, 2] = M[:, 0, 2] * 1.25 - 20.0
    M[:, 1, 2] = M[:, 1, 2] * 1.25 - 9.0
    ref_m, amb_m = _oracle_literal(c32, vw, c3, cube, M=M)
    img_d, juvd_d = rnd.getDepth(vw.cuda(), jw.cuda(), c3g, cubeg, M.cuda())
    _assert_same_image(img_d, ref_m, amb_m, "getDepth(M)")
    np.testing.assert_allclose(juvd_d.cpu().numpy(), mo.joint_trans(jw, M, c2, cube, NYU, 128).numpy(), rtol=0, atol=3e-4)
    rot = torch.tensor([[0.0, 0.7, 0.0]]).repeat(B, 1)
    from dsf_b200.mano_layer import batch_rodrigues
    Rm = batch_rodrigues(rot)
    vw_r = torch.einsum("bij,bvj->bvi", Rm, vw - c3[:, None]) + c3[:, None]
    ref_r, amb_r = _oracle_literal(c32, vw_r, c3, cube, M=M)
    img_r, _ = rnd.getDepth(vw.cuda(), jw.cuda(), c3g, cubeg, M.cuda(), rot.cuda())
    _assert_same_image(img_r, ref_r, amb_r, "getDepth(M, rot)")

    # M_render: raw MANO millimetres (no global_scale) placed by the cam parameters, cropped with M
    v_mm, _ = mo.get_mano_vertices(c32, q, t, b, cam)
    p_mm = p.clone()
    p_mm[:, 58] = 1.0
    p_mm[:, 59:62] = c3                                   # cam translation puts the hand at the crop centre
    qm, tm, bm, camm = mo.split_params(p_mm)
    v_mm, _ = mo.get_mano_vertices(c32, qm, tm, bm, camm)
    ref_mr, amb_mr = _oracle_literal(c32, v_mm, c3, cube, M=M)
    _assert_same_image(rnd.M_render(p_mm.cuda(), c3g, cubeg, M.cuda(), mask=False), ref_mr, amb_mr, "M_render")

    # forward (the synthetic-data generator): compare its image with the oracle on the placement it reports
    aug_view = torch.tensor([[0.3, -0.2, 0.5]]).repeat(B, 1)
    aug_shape = 0.5 * torch.ones(B, 10)
    out = rnd.forward(pg, c3g, cubeg, augmentView=aug_view.cuda(), augmentShape=aug_shape.cuda(),
                      augmentCenter=torch.tensor([[4.0, -3.0, 6.0]]).repeat(B, 1).cuda(),
                      augmentSize=torch.full((B, 3), 1.1).cuda(), mask=False)
    img_f, juvd_f, vuvd_f, jxyz_f, vxyz_f, c3_f, cube_f, M_f = out
    vw_f = (vxyz_f * cube_f[:, None] / 2 + c3_f[:, None]).cpu()
    ref_f, amb_f = _oracle_literal(c32, vw_f, c3_f.cpu(), cube_f.cpu())
    _assert_same_image(img_f, ref_f, amb_f, "forward")
    # ... and the placement itself against the reference's recipe in the oracle's MANO
    v_o, j_o = mo.get_mano_vertices(c32, q, t, b + aug_shape, cam)
    ctr = j_o.mean(1, keepdim=True)
    Ra = batch_rodrigues(aug_view)
    v_o = torch.einsum("bij,bvj->bvi", Ra, v_o - ctr) + c3[:, None]
    c3_o, cube_o = c3 + torch.tensor([4.0, -3.0, 6.0]), cube * 1.1
    np.testing.assert_allclose(c3_f.cpu().numpy(), c3_o.numpy(), rtol=1e-6)
    np.testing.assert_allclose(vxyz_f.cpu().numpy(), ((v_o - c3_o[:, None]) / cube_o[:, None] * 2).numpy(), rtol=0, atol=3e-5)


def test_direct_aligned_mode_registers_with_literal_chain(rnd):
    """the direct raster samples crop pixel i at i + 1/2, the literal chain (and M, JointTrans, the
    loader's crops) at i.  'direct_aligned' moves the principal point by half a crop pixel: its silhouettes sit on
    the literal ones with no systematic offset, the plain direct ones are half a pixel off."""
    from dsf_b200.fit import FitStep

    B, R = 64, 128
    inp = _inputs(B, seed=31)
    cen = {}
    for mode in ("direct", "direct_aligned", "literal"):
        st = FitStep(rnd.mano_layer, B, R, mode=mode, use_graph=False)
        st.set_inputs(inp["params"].cuda(), inp["center3d"].cuda(), inp["cube"].cuda())
        st.render_target(inp["params_target"].cuda())
        fg = (st.target < 0.99).float()
        n = fg.sum((1, 2)).clamp(min=1)
        ii = torch.arange(R, device="cuda", dtype=torch.float32)
        cen[mode] = torch.stack(((fg.sum(1) * ii).sum(1) / n, (fg.sum(2) * ii).sum(1) / n), 1).cpu()    # (col, row) centroid
        assert (n > 200).all()
    d_al = cen["direct_aligned"] - cen["literal"]
    d_di = cen["direct"] - cen["literal"]
    # (the literal chain's own 640 -> 480 nearest resize leaves ~0.08 px of vertical bias)
    assert d_al.abs().mean() < 0.15 and d_al.mean(0).abs().max() < 0.15, d_al.mean(0)
    assert 0.3 < d_di.mean(0).abs().min() and d_di.mean(0).abs().max() < 0.7, d_di.mean(0)
