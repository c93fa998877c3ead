"""GPU parity at the sizes and kernel variants bench.py actually runs (VERDICT r01 "next" item 1), and of the
rasteriser's fused loss-gradient epilogue.  Full-size batches run on the GPU; the CPU oracle checks a seeded
subset of hands (first / last hand of every slice included)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NYU = (588.03, 587.07, 320.0, 240.0)
RTOL_FWD = 1e-5
RTOL_GRAD = 1e-4


@pytest.fixture(scope="module")
def layer(mano_model):
    from dsf_b200.mano_layer import MANO_SMPL

    torch.cuda.set_device(0)
    return MANO_SMPL(mano_model, "nyu")


@pytest.fixture(scope="module")
def consts(mano_model):
    from oracle import mano_oracle as mo

    return mo.ManoConstants(mano_model), mo.ManoConstants(mano_model, torch.float64)


def _inputs(B, seed):
    from dsf_b200 import sample_fit_inputs

    return {k: torch.from_numpy(v) for k, v in sample_fit_inputs(B, seed=seed).items()}


def _subset(B, n, bounds, seed=0):
    g = torch.Generator().manual_seed(seed)
    must = sorted({i for lo, hi in bounds for i in (lo, hi - 1)})
    rest = torch.randperm(B, generator=g)[:n].tolist()
    idx = must + [i for i in rest if i not in must]
    return torch.tensor(idx[:max(n, len(must))])


def _view(mode, c3, cube, R):
    from dsf_b200 import _lib as L

    B = c3.shape[0]
    view = torch.empty(B, L.VIEW_STRIDE, device="cuda")
    xs, ys = torch.empty(B, R, device="cuda"), torch.empty(B, R, device="cuda")
    M = torch.empty(B, 3, 3, device="cuda")
    intr = (C.c_float * 4)(*NYU)
    L.check(L.lib().dsf_view_setup(mode, B, c3.data_ptr(), cube.data_ptr(), intr, 640, 480, R, None, view.data_ptr(),
                                   xs.data_ptr(), ys.data_ptr(), M.data_ptr(), L.stream_ptr()))
    return view, xs, ys, M


def _posed(layer, inp):
    p = inp["params"].cuda()
    v, j = layer.get_mano_vertices(p[:, :3], p[:, 3:48], p[:, 48:58], p[:, 58:], global_scale=1 / 125)
    c3, cube = inp["center3d"].cuda(), inp["cube"].cuda()
    return (v * cube[:, None] / 2 + c3[:, None]).contiguous().detach()


def _at(x_ref, x_gpu):
    """x_ref's autograd graph evaluated at exactly x_gpu (straight-through): the two MANO implementations differ
    by ~1e-7, and the loss gradient of a hand with a visible sliver face (gradient ~ 1 / area) reacts to that at
    the 1e-3 level, so the rasteriser is fed the very vertices the GPU produced; both subtractions are exact."""
    return x_ref + (x_gpu - x_ref).detach()


def _oracle_fit_grads(c32, prm, c3, cube, target, norm_batch, v_gpu, view8, xs, ys, pc=None):
    """per-hand loss terms and d(loss)/d(params) of the oracle chain (its own MANO autograd; rasterisation at the
    GPU's vertices) for a subset; loss = 0.1 / norm_batch * sum over hands."""
    from oracle import mano_oracle as mo
    from oracle import raster_oracle as ro

    p = prm.clone().requires_grad_(True)
    q, t, b, cam = mo.split_params(p)
    v_ref, _ = mo.get_mano_vertices(c32, q, t, b, cam, global_scale=1 / 125)
    vw = _at(v_ref, v_gpu) * cube[:, None] / 2 + c3[:, None]
    zbuf, p2f = ro.RasterDepth.apply(vw, c32.faces, view8, xs, ys, pc)
    img = ro.normalize_depth(zbuf, view8)
    mask = target.lt(0.99) | img.lt(0.99)
    per = (torch.abs(target - img) * mask).sum((-1, -2)) / (mask.float().sum((-1, -2)) + 1e-8)
    loss = per.sum() * 0.1 / norm_batch
    (g,) = torch.autograd.grad(loss, p)
    return per.detach(), g, img.detach(), p2f, v_ref.detach()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("R,B", [(128, 48), (256, 10), (132, 6)])
def test_fused_loss_gradient_equals_backward_kernel_and_oracle(layer, consts, mode, R, B):
    """dsf_raster_loss_grad (integer face moments + closed-form face gradient in the forward epilogue) against
    (a) the modular chain dsf_raster_forward -> dsf_depth_loss -> dsf_raster_backward on the same images and
    (b) the oracle's per-pixel backward.  R = 132 is not a multiple of 128/64: ragged tiles, scalar epilogue."""
    from dsf_b200 import _lib as L
    from oracle import raster_oracle as ro

    c32, _ = consts
    lib = L.lib()
    inp = _inputs(B, seed=300 + R + mode)
    vw = _posed(layer, inp)
    tgt_inp = dict(inp)
    tgt_inp["params"] = inp["params_target"]
    vw_t = _posed(layer, tgt_inp)
    c3, cube = inp["center3d"].cuda(), inp["cube"].cuda()
    view, xs, ys, _ = _view(mode, c3, cube, R)
    h, s = layer._handle, L.stream_ptr()
    f = lambda *sh: torch.empty(*sh, device="cuda")
    target, p2f_t = f(B, R, R), torch.empty(B, R, R, dtype=torch.int32, device="cuda")
    L.check(lib.dsf_raster_forward(h, B, vw_t.data_ptr(), view.data_ptr(), xs.data_ptr(), ys.data_ptr(), R,
                                   target.data_ptr(), p2f_t.data_ptr(), None, None, None, None, 0.0, None, 0, s))
    # fused
    img, p2f = f(B, R, R), torch.empty(B, R, R, dtype=torch.int32, device="cuda")
    parts, totals, gv = f(B, 2), f(4), f(B, 779, 3)
    ws = f(lib.dsf_raster_loss_workspace_floats(B, R))
    L.check(lib.dsf_raster_loss_grad(h, B, vw.data_ptr(), view.data_ptr(), xs.data_ptr(), ys.data_ptr(), R,
                                     target.data_ptr(), 0.99, 0.1, 0, img.data_ptr(), p2f.data_ptr(), parts.data_ptr(),
                                     totals.data_ptr(), gv.data_ptr(), ws.data_ptr(), 0, s))
    # the same call again without the pix_to_face plane: identical results (nothing reads it)
    img_b, parts_b, totals_b, gv_b = f(B, R, R), f(B, 2), f(4), f(B, 779, 3)
    L.check(lib.dsf_raster_loss_grad(h, B, vw.data_ptr(), view.data_ptr(), xs.data_ptr(), ys.data_ptr(), R,
                                     target.data_ptr(), 0.99, 0.1, 0, img_b.data_ptr(), None, parts_b.data_ptr(),
                                     totals_b.data_ptr(), gv_b.data_ptr(), ws.data_ptr(), 0, s))
    # modular chain
    img2, p2f2 = f(B, R, R), torch.empty(B, R, R, dtype=torch.int32, device="cuda")
    parts2, totals2, g_img, gv2 = f(B, 2), f(4), f(B, R, R), f(B, 779, 3)
    L.check(lib.dsf_raster_forward(h, B, vw.data_ptr(), view.data_ptr(), xs.data_ptr(), ys.data_ptr(), R,
                                   img2.data_ptr(), p2f2.data_ptr(), None, None, None, None, 0.0, None, 0, s))
    L.check(lib.dsf_depth_loss(0, B, R, target.data_ptr(), img2.data_ptr(), 0.99, 0.1, parts2.data_ptr(),
                               totals2.data_ptr(), g_img.data_ptr(), s))
    L.check(lib.dsf_raster_backward(h, B, vw.data_ptr(), view.data_ptr(), xs.data_ptr(), ys.data_ptr(), R,
                                    p2f2.data_ptr(), g_img.data_ptr(), gv2.data_ptr(), 0, s))
    torch.cuda.synchronize()
    if not torch.equal(img, img2):
        bad = (img != img2)
        w = bad.nonzero()
        raise AssertionError(f"{int(bad.sum())} pixels differ; first {w[:6].tolist()}; fused {img[bad][:6].tolist()} "
                             f"plain {img2[bad][:6].tolist()} p2f {p2f[bad][:6].tolist()} / {p2f2[bad][:6].tolist()}")
    assert torch.equal(p2f, p2f2)
    assert torch.equal(img_b, img)
    assert torch.equal(parts_b, parts) and torch.equal(totals_b, totals)
    torch.testing.assert_close(gv_b, gv, rtol=1e-5, atol=1e-9)       # float atomics on shared memory reorder sums
    assert (p2f >= 0).float().mean() > 0.03
    torch.testing.assert_close(parts, parts2, rtol=2e-6, atol=1e-6)
    torch.testing.assert_close(totals, totals2, rtol=1e-5, atol=1e-7)
    # yardstick: the oracle's backward in float64 on the GPU's own pix_to_face with the same cotangent (the
    # float32 per-pixel sums - reference, oracle and raster_bwd_kernel alike - carry cancellation noise of their own
    # on sliver faces, so they are compared with the float64 result, not with each other)
    v8 = view[:, :8].cpu().contiguous()
    p_ref, z_ref, _, vndc = ro.render(vw.cpu(), c32.faces, v8, xs.cpu(), ys.cpu(), perspective_correct=False)
    assert torch.equal(p_ref, p2f.cpu())
    zc, zh = v8[:, 4].view(-1, 1, 1), v8[:, 5].view(-1, 1, 1)
    live = (p_ref >= 0) & (z_ref > 0) & (z_ref <= zc + zh) & (z_ref >= zc - zh)
    g_z = torch.where(live, g_img.cpu() / zh, torch.zeros(1))
    gv64 = ro.render_backward_f64(vw.cpu(), c32.faces, v8, xs.cpu(), ys.cpu(), p_ref, g_z, perspective_correct=False,
                                  vndc=vndc)
    scale = gv64.abs().amax((1, 2)).clamp(min=1e-30)
    err_fused = (gv.cpu().double() - gv64).abs().amax((1, 2)) / scale
    err_mod = (gv2.cpu().double() - gv64).abs().amax((1, 2)) / scale
    gv32 = ro.render_backward(vw.cpu(), c32.faces, v8, xs.cpu(), ys.cpu(), p_ref, g_z, vndc, perspective_correct=False)
    err_o32 = (gv32.double() - gv64).abs().amax((1, 2)) / scale
    print(f"[fused grad R={R} mode={mode}] max rel err vs f64: fused {err_fused.max():.2e}, raster_bwd_kernel "
          f"{err_mod.max():.2e}, f32 oracle {err_o32.max():.2e}")
    # both kernels: 1e-4 of the float64 result (the float32 oracle, like the reference, sits ~1e-3 away).  The
    # integer-moment epilogue idealises the sample positions as 1 - (2 q + 1) / S, exact for the direct raster; in
    # the literal 640-pixel raster the float32 coordinates deviate by up to an ulp (2e-5 of a pixel) and sliver
    # faces amplify that to ~1e-3, which is why FitStep keeps the per-pixel kernel there (DSF_RASTER_SEPARATE_BACKWARD)
    assert err_fused.max() < (RTOL_GRAD if mode == 0 else 2e-3), err_fused
    assert err_mod.max() < RTOL_GRAD, err_mod


def test_bench_size_fit_step_4096_subset_vs_oracle(layer, consts):
    """The configuration bench.py times - FitStep(B=4096, two stream slices, CUDA graph replay) - against the
    oracle on 128 seeded hands incl. the first / last hand of each slice: pix_to_face and the image bit-exact
    (oracle rasteriser on the GPU's own vertices), vertices 1e-5, per-hand loss 1e-5, parameter gradients 1e-4 on
    the hands that carry no rounding-decided pixel (their count is asserted)."""
    from dsf_b200.fit import FitStep
    from oracle import mano_oracle as mo
    from oracle import raster_oracle as ro

    c32, _ = consts
    B, R = 4096, 128
    inp = _inputs(B, seed=1000)
    step = FitStep(layer, B, R, use_graph=True, chunks=2)
    step.set_inputs(inp["params"].cuda(), inp["center3d"].cuda(), inp["cube"].cuda())
    step.render_target(inp["params_target"].cuda())
    step.step()
    step.step()
    torch.cuda.synchronize()
    idx = _subset(B, 128, step._bounds, seed=1)
    n = len(idx)
    ic = idx.cuda()
    prm, c3, cube = inp["params"][idx], inp["center3d"][idx], inp["cube"][idx]
    target = step.target[ic].cpu()
    v_gpu = step.verts[ic].cpu()
    v8 = step.view[ic, :8].cpu().contiguous()
    per_ref, g_ref, img_ref, p2f_ref, v_ref = _oracle_fit_grads(c32, prm, c3, cube, target, B, v_gpu, v8,
                                                                  step.xs[ic].cpu(), step.ys[ic].cpu())
    assert ((v_gpu - v_ref).abs().amax((1, 2)) / v_ref.abs().amax((1, 2))).max() < RTOL_FWD
    # forward products of every compared hand: bit-exact
    assert torch.equal(step.p2f[ic].cpu(), p2f_ref)
    assert torch.equal(step.img[ic].cpu(), img_ref)
    parts = step.parts[ic].cpu()
    per_gpu = parts[:, 0] / (parts[:, 1] + 1e-8)
    assert ((per_gpu - per_ref).abs() / per_ref.abs().clamp(min=1e-12)).max() < RTOL_FWD
    # parameter gradients of every compared hand: nothing excluded (identical images leave no rounding-decided pixel)
    g = step.g_params[ic].cpu()
    per_hand = (g - g_ref).abs().amax(1) / g_ref.abs().amax(1)
    print(f"[bench-size parity] {n} hands, 0 excluded; max rel grad err {per_hand.max():.2e}")
    assert per_hand.max() < RTOL_GRAD, per_hand
    # whole-batch loss = mean of the per-hand terms
    full = step.parts[:, 0] / (step.parts[:, 1] + 1e-8)
    assert abs(step.totals[0].item() - 0.1 * full.double().mean().item()) < 1e-5 * abs(step.totals[0].item())


@pytest.mark.parametrize("pc", [False, True])
def test_raster_backward_large_batch_variant_vs_oracle(layer, consts, pc):
    """n_mesh >= 2048 selects raster_bwd_kernel<256> (8192-pixel chunks, two passes at R = 128), the variant
    the perspective-correct fused step and the modular path use at bench sizes."""
    from dsf_b200 import _lib as L
    from oracle import raster_oracle as ro

    c32, _ = consts
    lib = L.lib()
    B, R = 2048, 128
    inp = _inputs(B, seed=55)
    vw = _posed(layer, inp)
    c3, cube = inp["center3d"].cuda(), inp["cube"].cuda()
    view, xs, ys, _ = _view(0, c3, cube, R)
    img = torch.empty(B, R, R, device="cuda")
    p2f = torch.empty(B, R, R, dtype=torch.int32, device="cuda")
    s = L.stream_ptr()
    L.check(lib.dsf_raster_forward(layer._handle, B, vw.data_ptr(), view.data_ptr(), xs.data_ptr(), ys.data_ptr(), R,
                                   img.data_ptr(), p2f.data_ptr(), None, None, None, None, 0.0, None, int(pc), s))
    g_img = torch.randn(B, R, R, generator=torch.Generator().manual_seed(4)).cuda()
    gv = torch.empty_like(vw)
    L.check(lib.dsf_raster_backward(layer._handle, B, vw.data_ptr(), view.data_ptr(), xs.data_ptr(), ys.data_ptr(), R,
                                    p2f.data_ptr(), g_img.data_ptr(), gv.data_ptr(), int(pc), s))
    torch.cuda.synchronize()
    idx = _subset(B, 48, [(0, B)], seed=2)
    ic = idx.cuda()
    v8 = view[ic, :8].cpu().contiguous()
    p_ref, z_ref, _, vndc = ro.render(vw[ic].cpu(), c32.faces, v8, xs[ic].cpu(), ys[ic].cpu(), perspective_correct=pc)
    assert torch.equal(p_ref, p2f[ic].cpu())
    zc, zh = v8[:, 4].view(-1, 1, 1), v8[:, 5].view(-1, 1, 1)
    live = (p_ref >= 0) & (z_ref > 0) & (z_ref <= zc + zh) & (z_ref >= zc - zh)
    g_z = torch.where(live, g_img[ic].cpu() / zh, torch.zeros(1))
    gv_ref = ro.render_backward_f64(vw[ic].cpu(), c32.faces, v8, xs[ic].cpu(), ys[ic].cpu(), p_ref, g_z,
                                    perspective_correct=pc, vndc=vndc)  # float64 yardstick, see test_gpu_parity
    per_hand = (gv[ic].cpu().double() - gv_ref).abs().amax((1, 2)) / gv_ref.abs().amax((1, 2))
    assert per_hand.max() < RTOL_GRAD, per_hand


@pytest.mark.parametrize("B", [2048, 4096])
def test_mano_backward_bench_batch_vs_oracle(layer, consts, B):
    """MANO forward + backward at the bench's slice sizes (the split-K factor of the transposed blend GEMM
    depends on the batch: 9-way at 2048 hands) against float64 autograd of the oracle on a subset."""
    from oracle import mano_oracle as mo

    _, c64 = consts
    inp = _inputs(B, seed=7)
    p = inp["params"].cuda().requires_grad_(True)
    v, j = layer.get_mano_vertices(p[:, :3], p[:, 3:48], p[:, 48:58], p[:, 58:], global_scale=1 / 125)
    gen = torch.Generator().manual_seed(3)
    wv, wj = torch.randn(B, 779, 3, generator=gen), torch.randn(B, 21, 3, generator=gen)
    ((v * wv.cuda()).sum() + (j * wj.cuda()).sum()).backward()
    torch.cuda.synchronize()
    idx = _subset(B, 64, [(0, B)], seed=5)
    pr = inp["params"][idx].double().requires_grad_(True)
    q, t, b, cam = mo.split_params(pr)
    v_ref, j_ref = mo.get_mano_vertices(c64, q, t, b, cam, global_scale=1 / 125)
    ((v_ref * wv[idx].double()).sum() + (j_ref * wj[idx].double()).sum()).backward()
    assert ((v[idx.cuda()].detach().cpu().double() - v_ref.detach()).abs().amax((1, 2)) / v_ref.detach().abs().amax((1, 2))).max() < RTOL_FWD
    g, g_ref = p.grad[idx.cuda()].cpu().double(), pr.grad
    per_hand = (g - g_ref).abs().amax(1) / g_ref.abs().amax(1)
    assert per_hand.max() < RTOL_GRAD, per_hand


def test_icp_config4_full_size_subset_vs_oracle(layer, consts):
    """BASELINE configs[4]: ICPLoss 1024 hands x 2048 points x 1554 faces; distances / arg-min / gradients of a
    32-hand subset against the C oracle."""
    from dsf_b200.mesh_loss import _PointFaceDistance
    from oracle import raster_oracle as ro

    c32, _ = consts
    B, P = 1024, 2048
    inp = _inputs(B, seed=5)
    p = inp["params"].cuda()
    v, _ = layer.get_mano_vertices(p[:, :3], p[:, 3:48], p[:, 48:58], p[:, 58:], global_scale=1 / 125)
    v = v.detach()
    gen = torch.Generator(device="cuda").manual_seed(0)
    pidx = torch.randint(0, 778, (B, P), device="cuda", generator=gen)
    pcl = torch.gather(v, 1, pidx[..., None].expand(-1, -1, 3)) + 0.05 * torch.randn(B, P, 3, device="cuda", generator=gen)
    vg, pg = v.clone().requires_grad_(True), pcl.clone().requires_grad_(True)
    d, fi = _PointFaceDistance.apply(pg, vg, layer.faces_int)
    w = torch.rand(B, P, device="cuda", generator=gen)
    gp, gvv = torch.autograd.grad((d * w).sum(), (pg, vg))
    torch.cuda.synchronize()
    idx = _subset(B, 32, [(0, B)], seed=9)
    ic = idx.cuda()
    d_ref, i_ref = ro.point_face(pcl[ic].cpu(), v[ic].cpu(), c32.faces)
    dg = d[ic].detach().cpu()
    assert (np.abs(dg - d_ref) <= RTOL_FWD * d_ref.abs().max(1, keepdim=True)[0] + 1e-12).all()
    same = fi[ic].cpu() == i_ref
    n_tie = int((~same).sum())
    assert n_tie <= 0.1 * same.numel(), n_tie
    assert ((dg - d_ref).abs()[~same] <= 1e-5 * d_ref[~same] + 1e-9).all()      # ties: equidistant faces
    gp_ref, gv_ref = ro.point_face_backward(pcl[ic].cpu(), v[ic].cpu(), c32.faces, fi[ic].cpu(), w[ic].cpu(), double=True)
    assert (gp[ic].cpu().double() - gp_ref).abs().max() / gp_ref.abs().max() < RTOL_GRAD
    assert (gvv[ic].cpu().double() - gv_ref).abs().max() / gv_ref.abs().max() < RTOL_GRAD
    print(f"[icp full size] arg-min ties {n_tie} of {same.numel()}")


@pytest.mark.parametrize("pc", [False, True])
def test_multiview_config3_full_size_subset_vs_oracle(layer, consts, pc):
    """BASELINE configs[3]: 512 hands x 3 views x 256^2 through the fused multi-view step; a subset of hands
    against the oracle: per-view pix_to_face / image bit-exact on the GPU's own rotated vertices, and the summed
    parameter gradient against autograd through the oracle chain (MANO -> rotation -> raster -> m2d loss)."""
    from dsf_b200.fit import MultiViewFitStep
    from dsf_b200.mano_layer import batch_rodrigues
    from oracle import mano_oracle as mo
    from oracle import raster_oracle as ro

    c32, _ = consts
    B, V, R = 512, 3, 256
    inp = _inputs(B, seed=9)
    rot = torch.tensor([[0.0, 0.0, 0.0], [0.0, 2 * np.pi / 3, 0.0], [0.0, -2 * np.pi / 3, 0.0]]).repeat(B, 1).reshape(B, V, 3)
    mv = MultiViewFitStep(layer, B, V, R, use_graph=True, perspective_correct=pc)
    mv.set_inputs(inp["params"].cuda(), inp["center3d"].cuda(), inp["cube"].cuda(), rot.cuda())
    mv.render_target(inp["params_target"].cuda())
    mv.step()
    mv.step()
    torch.cuda.synchronize()
    idx = _subset(B, 6, [(0, B)], seed=4)
    n = len(idx)
    midx = (idx[:, None] * V + torch.arange(V)[None]).flatten()          # mesh rows of the subset, view-minor
    mc = midx.cuda()
    Rm = batch_rodrigues(rot[idx].reshape(-1, 3)).reshape(n, V, 3, 3)
    c3, cube = inp["center3d"][idx], inp["cube"][idx]
    v8 = mv.view[mc, :8].cpu().contiguous()
    xs, ys = mv.xs[mc].cpu(), mv.ys[mc].cpu()
    target = mv.target[mc].cpu()

    def place(v):                      # (n,779,3) normalised -> (n*V,779,3) camera space, rotated about the centre
        x = v * cube[:, None] / 2
        return (torch.einsum("nvij,nkj->nvki", Rm.to(v.dtype), x) + c3[:, None, None]).reshape(n * V, 779, 3)

    vw_gpu = mv.verts_cam[mc].cpu()                     # the rotated vertices the rasteriser really saw
    assert ((vw_gpu - place(mv.verts[idx.cuda()].cpu())).abs().max() / vw_gpu.abs().max()) < 1e-6
    p_ref, z_ref, _, _ = ro.render(vw_gpu, c32.faces, v8, xs, ys, perspective_correct=pc)
    assert torch.equal(mv.p2f[mc].cpu(), p_ref), int((mv.p2f[mc].cpu() != p_ref).sum())
    assert torch.equal(mv.img[mc].cpu(), ro.normalize_depth(z_ref, v8))
    assert (p_ref >= 0).float().mean() > 0.03
    # gradients: oracle autograd chain on the subset (MANO -> rotation -> raster at the GPU's own rotated vertices ->
    # m2d loss); loss = 0.1 / (B V) * sum over meshes
    pr = inp["params"][idx].clone().requires_grad_(True)
    q, t, b, cam = mo.split_params(pr)
    v_ref, _ = mo.get_mano_vertices(c32, q, t, b, cam, global_scale=1 / 125)
    z, p2f_o = ro.RasterDepth.apply(_at(place(v_ref), vw_gpu), c32.faces, v8, xs, ys, pc)
    assert torch.equal(p2f_o, p_ref)
    img_o = ro.normalize_depth(z, v8)
    mask = target.lt(0.99) | img_o.lt(0.99)
    per = (torch.abs(target - img_o) * mask).sum((-1, -2)) / (mask.float().sum((-1, -2)) + 1e-8)
    (g_ref,) = torch.autograd.grad(per.sum() * 0.1 / (B * V), pr)
    g = mv.g_params[idx.cuda()].cpu()
    per_hand = (g - g_ref).abs().amax(1) / g_ref.abs().amax(1)
    print(f"[C3 full size pc={pc}] {n} hands, 0 excluded; rel err {per_hand}")
    assert per_hand.max() < RTOL_GRAD, per_hand
