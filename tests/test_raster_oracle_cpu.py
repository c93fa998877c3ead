"""Self-checks of the C oracle for the un-vendored pytorch3d arithmetic (parity unpinned by the
reference, so: known answers, float64 cross-evaluation, finite differences, brute force)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import mano_oracle as mo
from oracle import raster_oracle as ro

NYU = (588.03, 587.07, 320.0, 240.0)


def _scene(mano_model, B=3, seed=0):
    from dsf_b200 import sample_fit_inputs

    c = mo.ManoConstants(mano_model)
    inp = {k: torch.from_numpy(v) for k, v in sample_fit_inputs(B, seed=seed).items()}
    q, t, b, cam = mo.split_params(inp["params"])
    v, j = mo.get_mano_vertices(c, q, t, b, cam, global_scale=1 / 125)
    vw = v * inp["cube"][:, None] / 2 + inp["center3d"][:, None]
    return c, inp, vw


def test_single_triangle_known_answer():
    # a fronto-parallel triangle at z=1000 covering the image centre: depth exactly 1000 inside
    verts = torch.tensor([[[-50.0, -50.0, 1000.0], [60.0, -40.0, 1000.0], [0.0, 70.0, 1000.0]]])
    faces = torch.tensor([[0, 1, 2]])
    view = torch.zeros(1, 8)
    view[0, :4] = torch.tensor([2 * 588.03 / 128, 2 * 587.07 / 128, 0.0, 0.0])
    view[0, 4], view[0, 5] = 1000.0, 125.0
    xs = ro.pix_to_ndc(128)[None]
    p2f, z, bary, _ = ro.render(verts, faces, view, xs, xs.clone(), want_bary=True)
    fg = p2f[0] >= 0
    assert fg.sum() > 500 and fg[64, 64]
    assert torch.allclose(z[0][fg], torch.tensor(1000.0), rtol=1e-6)
    assert torch.allclose(bary[0][fg].sum(-1), torch.tensor(1.0), atol=1e-5)
    assert (z[0][~fg] == -1).all() and (p2f[0][~fg] == -1).all()
    # image convention: +x (camera right) appears at larger column index, +y (down) at larger row
    cols = torch.nonzero(fg.any(0)).flatten()
    rows = torch.nonzero(fg.any(1)).flatten()
    assert cols.max() - 64 > 64 - cols.min()        # triangle extends further to +x
    assert rows.max() - 64 > 64 - rows.min()        # and further to +y


def test_lowest_face_index_wins_exact_ties():
    verts = torch.tensor([[[-50.0, -50.0, 900.0], [60.0, -40.0, 900.0], [0.0, 70.0, 900.0]]])
    faces = torch.tensor([[0, 1, 2], [0, 1, 2], [0, 1, 2]])
    view = torch.zeros(1, 8)
    view[0, :4] = torch.tensor([9.0, 9.0, 0.0, 0.0])
    view[0, 4], view[0, 5] = 900.0, 125.0
    xs = ro.pix_to_ndc(64)[None]
    p2f, _, _, _ = ro.render(verts, faces, view, xs, xs.clone())
    assert set(p2f.unique().tolist()) == {-1, 0}


@pytest.mark.parametrize("pc", [False, True])
def test_f32_and_f64_rasterisers_agree_except_rounding_ties(mano_model, pc):
    c, inp, vw = _scene(mano_model, 4, 3)
    for mode in ("direct", "literal"):
        view, xs, ys, _ = ro.make_view(mode, inp["center3d"], inp["cube"], NYU, 640, 480, 128)
        p32, z32, _, _ = ro.render(vw, c.faces, view, xs, ys, perspective_correct=pc)
        p64, z64 = ro.render_f64(vw, c.faces, view, xs, ys, perspective_correct=pc)
        assert (p32 >= 0).float().mean() > 0.05
        assert (p32 != p64).float().mean() < 1e-3
        same = (p32 == p64) & (p32 >= 0)
        assert ((z32[same].double() - z64[same]).abs() / z64[same]).max() < 1e-5


@pytest.mark.parametrize("pc", [0, 1])
def test_raster_backward_matches_finite_differences(mano_model, pc):
    c, inp, vw = _scene(mano_model, 1, 5)
    view, xs, ys, _ = ro.make_view("direct", inp["center3d"], inp["cube"], NYU, 640, 480, 128)
    L = ro.lib()
    V = vw.shape[1]
    faces = c.faces.int().contiguous()
    R = 128
    xs64, ys64 = xs[0].double().contiguous(), ys[0].double().contiguous()
    cf, cd, ci = ctypes.c_float, ctypes.c_double, ctypes.c_int
    v4 = [cf(float(view[0, i])) for i in range(4)]

    def fwd(verts32):
        vn = torch.empty(V, 3, dtype=torch.float64)
        L.orc_project_f64(ro._p(verts32), V, *v4, ro._p(vn, cd))
        p = torch.empty(R, R, dtype=torch.int32)
        z = torch.empty(R, R, dtype=torch.float64)
        L.orc_rasterize_f64(ro._p(vn, cd), ro._p(faces, ci), faces.shape[0], ro._p(xs64, cd), R, ro._p(ys64, cd),
                            R, pc, cd(1e-8), 0, ro._p(p, ci), ro._p(z, cd), None, None)
        return vn, p, z

    base = vw[0].contiguous()
    vn, p0, z0 = fwd(base)
    gz = torch.randn(R, R, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    gvn = torch.zeros(V, 3, dtype=torch.float64)
    L.orc_rasterize_backward_f64(ro._p(vn, cd), ro._p(faces, ci), ro._p(xs64, cd), R, ro._p(ys64, cd), R,
                                 ro._p(p0, ci), ro._p(gz, cd), None, pc, cd(1e-8), ro._p(gvn, cd))
    gv = torch.zeros(V, 3, dtype=torch.float64)
    L.orc_project_backward_f64(ro._p(base), V, *v4, ro._p(gvn, cd), ro._p(gv, cd))
    # finite differences on a few vertices that own foreground pixels (visibility held fixed by
    # keeping only pixels whose face does not change)
    used = torch.unique(faces[p0[p0 >= 0].long()].flatten())[:6]
    h = 2.0 ** -10     # exactly representable step (verts are float32, |x| < 8192 mm)
    for v in used.tolist():
        for ax in range(3):
            vp, vm = base.clone(), base.clone()
            vp[v, ax] += h
            vm[v, ax] -= h
            _, pp, zp = fwd(vp)
            _, pm, zm = fwd(vm)
            keep = (pp == p0) & (pm == p0) & (p0 >= 0)
            fd = ((zp - zm)[keep] * gz[keep]).sum() / (2 * h)
            # analytic gradient restricted to the same pixels
            gvn2 = torch.zeros(V, 3, dtype=torch.float64)
            pk = torch.where(keep, p0, torch.full_like(p0, -1)).contiguous()
            L.orc_rasterize_backward_f64(ro._p(vn, cd), ro._p(faces, ci), ro._p(xs64, cd), R, ro._p(ys64, cd), R,
                                         ro._p(pk, ci), ro._p(gz, cd), None, pc, cd(1e-8), ro._p(gvn2, cd))
            gv2 = torch.zeros(V, 3, dtype=torch.float64)
            L.orc_project_backward_f64(ro._p(base), V, *v4, ro._p(gvn2, cd), ro._p(gv2, cd))
            assert abs(fd - gv2[v, ax]) <= 2e-5 * max(1.0, abs(gv2[v, ax])), (v, ax, fd, gv2[v, ax])
    # float32 backward agrees with the float64 one
    p32, z32, _, vndc = ro.render(vw, c.faces, view, xs, ys, perspective_correct=bool(pc))
    g32 = ro.render_backward(vw, c.faces, view, xs, ys, p32, gz.float()[None], vndc, perspective_correct=bool(pc))
    if torch.equal(p32[0], p0):
        assert (g32[0].double() - gv).abs().max() <= 2e-4 * gv.abs().max()


def test_point_face_against_brute_force(mano_model):
    c, inp, vw = _scene(mano_model, 2, 9)
    g = torch.Generator().manual_seed(0)
    pts = vw[:, torch.randint(0, 778, (300,), generator=g)] + torch.randn(2, 300, 3, generator=g) * 8.0
    d, idx = ro.point_face(pts, vw, c.faces)
    # brute force in float64: dense sampling of barycentric coordinates bounds the true distance
    tri = vw[0].double()[c.faces]                     # (F,3,3)
    u = torch.linspace(0, 1, 25, dtype=torch.float64)
    a, b = torch.meshgrid(u, u, indexing="ij")
    keep = a + b <= 1
    w = torch.stack([a[keep], b[keep], 1 - a[keep] - b[keep]], -1)           # (S,3)
    samples = torch.einsum("sk,fkc->fsc", w, tri).reshape(-1, 3)
    spacing = (tri - tri.roll(1, 1)).norm(dim=-1).max() / 24
    for i in range(0, 40):
        dd = ((samples - pts[0, i].double()) ** 2).sum(-1).min()
        assert d[0, i] <= dd * (1 + 1e-6) + 1e-9
        assert d[0, i].sqrt() >= dd.sqrt() - spacing  # the sampling error of the brute force is bounded
    # gradient: finite differences of the f64 oracle on the points
    gp, gv = ro.point_face_backward(pts, vw, c.faces, idx, torch.ones(2, 300), double=True)
    L = ro.lib()
    faces = c.faces.int().contiguous()
    cd, ci, cf = ctypes.c_double, ctypes.c_int, ctypes.c_float

    def dist64(p):
        dists = torch.empty(1, dtype=torch.float64)
        ii = torch.empty(1, dtype=torch.int32)
        L.orc_point_face_forward_f64(ro._p(p.contiguous()), 1, ro._p(vw[0].contiguous()), ro._p(faces, ci),
                                     faces.shape[0], cd(1e-8), ro._p(dists, cd), ro._p(ii, ci))
        return dists.item()

    h = 2.0 ** -7
    for i in range(5):
        for ax in range(3):
            pp, pm = pts[0, i:i + 1].clone(), pts[0, i:i + 1].clone()
            pp[0, ax] += h
            pm[0, ax] -= h
            fd = (dist64(pp) - dist64(pm)) / (2 * h)
            assert abs(fd - gp[0, i, ax].item()) <= 1e-3 * max(1.0, abs(fd))


def test_point_face_sphere_bound_never_exceeds_reported_distance():
    """The cull of dsf_point_face_forward skips a face when |p - c| >= sqrt(best) + r, with the sphere built
    over the triangle scaled about v0 by (den + eps) / den (the region where the eps-regularised inside test of
    the reference returns the plane distance).  Checked here against the oracle's reported distance on random,
    small and ill-conditioned triangles: (|p - c| - r)^2 must never exceed it (0.02 % slack as in the kernel)."""
    from oracle import raster_oracle as ro

    gen = torch.Generator().manual_seed(11)
    T, P = 4000, 64
    scale = 10 ** (torch.rand(T, 1, 1, generator=gen) * 3 - 3)               # edge lengths 1e-3 .. 1
    tri = torch.randn(T, 3, 3, generator=gen) * scale
    tri[::7, 2] = tri[::7, 0] + (tri[::7, 1] - tri[::7, 0]) * 0.4 + 1e-6 * torch.randn(len(tri[::7]), 3, generator=gen)  # slivers
    tri = tri + torch.randn(T, 1, 3, generator=gen)
    pts = tri.mean(1, keepdim=True) + torch.randn(T, P, 3, generator=gen) * scale * 10 ** (torch.rand(T, P, 1, generator=gen) * 3 - 2)
    faces = torch.tensor([[0, 1, 2]], dtype=torch.int32)
    d, _ = ro.point_face(pts.contiguous(), tri.contiguous(), faces)
    v0, e1, e2 = tri[:, 0], tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    d00, d01, d11 = (e1 * e1).sum(-1), (e1 * e2).sum(-1), (e2 * e2).sum(-1)
    den = d00 * d11 - d01 * d01
    k = torch.where((den > 1e-4 * d00 * d11) & (den > 1e-30), 1.01 * (den + 1e-8) / den, torch.full_like(den, float("inf")))
    s1, s2 = e1 * k[:, None], e2 * k[:, None]
    cr = (s1 + s2) / 3
    rr = torch.stack([cr.norm(dim=-1), (cr - s1).norm(dim=-1), (cr - s2).norm(dim=-1)]).max(0)[0]
    ok = (k < 1e6) & (rr < 1e18)
    assert ok.float().mean() > 0.6 and (~ok).sum() > 10                        # both branches are exercised
    dist_c = (pts - (v0 + cr)[:, None]).norm(dim=-1)
    lb = (dist_c - rr[:, None]).clamp(min=0) ** 2
    viol = (lb > d * 1.0002 + 1e-30) & ok[:, None]
    assert viol.sum() == 0, (viol.sum(), (lb / d)[viol].max())
    # and the bound is useful: it exceeds a tenth of the distance for most far points
    far = ok[:, None] & (dist_c > 4 * rr[:, None])
    assert (lb[far] > 0.1 * d[far]).float().mean() > 0.9


def test_point_face_box_bound_never_exceeds_reported_distance():
    """The box hierarchy of the whole-mesh point-face kernel (point_face_fwd_all_kernel) skips a group of faces when
    the squared distance from the point to the group's axis-aligned box exceeds `best` (0.02 % slack).  The box of a
    face is the box of its triangle SCALED about v0 by k = 1.01 (den + eps) / den - the same region the bounding
    sphere covers - grown by 1e-6 of the largest coordinate; a group's box is the union.  Restated here in float32
    torch and checked against the oracle's reported distance on random, small and ill-conditioned triangles, alone
    and in groups of 8: the bound must never exceed the distance to any member."""
    from oracle import raster_oracle as ro

    gen = torch.Generator().manual_seed(5)
    T, P = 4000, 64
    scale = 10 ** (torch.rand(T, 1, 1, generator=gen) * 3 - 3)
    tri = torch.randn(T, 3, 3, generator=gen) * scale
    tri[::7, 2] = tri[::7, 0] + (tri[::7, 1] - tri[::7, 0]) * 0.4 + 1e-6 * torch.randn(len(tri[::7]), 3, generator=gen)
    tri = tri + torch.randn(T, 1, 3, generator=gen)
    pts = tri.mean(1, keepdim=True) + torch.randn(T, P, 3, generator=gen) * scale * 10 ** (torch.rand(T, P, 1, generator=gen) * 3 - 2)
    faces = torch.tensor([[0, 1, 2]], dtype=torch.int32)
    d, _ = ro.point_face(pts.contiguous(), tri.contiguous(), faces)          # (T, P) squared distances
    v0, e1, e2 = tri[:, 0], tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    d00, d01, d11 = (e1 * e1).sum(-1), (e1 * e2).sum(-1), (e2 * e2).sum(-1)
    den = d00 * d11 - d01 * d01
    k = torch.where((den > 1e-4 * d00 * d11) & (den > 1e-30), 1.01 * (den + 1e-8) / den, torch.full_like(den, float("inf")))
    ok = k < 1e6                                                               # the kernel's never-cull flag otherwise
    kk = torch.where(ok, k, torch.ones_like(k))
    s1, s2 = e1 * kk[:, None], e2 * kk[:, None]
    zero = torch.zeros_like(s1)
    lo = v0 + torch.minimum(zero, torch.minimum(s1, s2))
    hi = v0 + torch.maximum(zero, torch.maximum(s1, s2))
    inf = torch.full_like(lo, float("inf"))
    lo, hi = torch.where(ok[:, None], lo, -inf), torch.where(ok[:, None], hi, inf)

    def grow(lo, hi):
        m = 1e-6 * torch.maximum(lo.abs(), hi.abs()) + 1e-12
        return lo - m, hi + m

    def box_d2(p, lo, hi):                                                     # p (T,P,3), boxes (T,3)
        dd = torch.maximum(torch.maximum(lo[:, None] - p, p - hi[:, None]), torch.zeros_like(p))
        return (dd * dd).sum(-1)

    glo, ghi = grow(lo, hi)
    lb = box_d2(pts, glo, ghi)
    viol = lb > d * 1.0002 + 1e-30
    assert viol.sum() == 0, (int(viol.sum()), float((lb / d)[viol].max()))
    assert ok.float().mean() > 0.6 and (~ok).sum() > 10
    far = ok[:, None] & (lb > 0)
    assert (lb[far] > 0.1 * d[far]).float().mean() > 0.5                       # and it is a useful bound
    # groups of 8 consecutive faces: the union box bounds the distance of the group's points to EVERY member
    G = T // 8
    ulo, uhi = grow(lo[: G * 8].view(G, 8, 3).min(1)[0], hi[: G * 8].view(G, 8, 3).max(1)[0])
    pg = pts[: G * 8: 8]                                                       # the first member's points
    tri_g = tri[: G * 8].view(G, 8, 3, 3)
    lbg = box_d2(pg, ulo, uhi)
    for m in range(8):
        dm, _ = ro.point_face(pg.contiguous(), tri_g[:, m].contiguous(), faces)
        viol = lbg > dm * 1.0002 + 1e-30
        assert viol.sum() == 0, (m, int(viol.sum()))
