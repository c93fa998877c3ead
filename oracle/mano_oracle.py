"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's MANO path.

A plain-torch (CPU, any float dtype, autograd-capable) restatement of
/root/reference/render_model/mano_layer.py for: the constant set-up (M0, :98-149),
MANO_SMPL.forward (M1, :573-641), batch_rodrigues / quat2mat (M2, :697-728),
batch_global_rigid_transformation (M3, :730-770), get_mano_vertices (M4, :643-693),
the 66-sphere collision term (C1, :229-317, :373-385) and the pure-torch pieces of
Render (R3/R4/R5, :1071-1097, :1133-1169, :1233-1260, :1289-1324).

Pinned: tests/test_oracle_cpu.py checks this file against golden vectors produced by
the reference's own MANO_SMPL / Render helpers imported unmodified
(tests/golden/make_golden.py).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product never does.
"""
from __future__ import annotations

import numpy as np
import torch

PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
WRIST_RING = [121, 214, 215, 279, 239, 234, 92, 38, 122, 118, 117, 119, 120, 108, 79, 78]
TIP_VERTS = [333, 444, 672, 555, 744]
CHILD = [2, 3, 16, 5, 6, 17, 8, 9, 18, 11, 12, 19, 14, 15, 20]      # mano_layer.py:229


class ManoConstants:
    """Constant buffers in the layouts the reference builds (mano_layer.py:112-149)."""

    def __init__(self, model: dict, dtype=torch.float32):
        f = np.asarray(model["f"]).astype(np.int64)
        fan = [[WRIST_RING[i], WRIST_RING[(i + 1) % 16], 778] for i in range(16)]   # :103-105
        self.faces = torch.from_numpy(np.concatenate([f, np.asarray(fan)], 0))       # (1554,3)
        self.v_template = torch.tensor(np.asarray(model["v_template"]), dtype=dtype)  # (778,3)
        sd = np.asarray(model["shapedirs"])
        self.shapedirs = torch.tensor(sd.reshape(-1, sd.shape[-1]).T.copy(), dtype=dtype)   # (10,2334) :118
        pd = np.asarray(model["posedirs"])
        self.posedirs = torch.tensor(pd.reshape(-1, pd.shape[-1]).T.copy(), dtype=dtype)    # (135,2334) :144
        jr = model["J_regressor"]
        jr = jr.toarray() if hasattr(jr, "toarray") else np.asarray(jr)
        jr = jr.T                                                                    # (778,16) :123
        tips = np.zeros((778, 5))
        for c, v in enumerate(TIP_VERTS):                                            # :124-131
            tips[v, c] = 1.0
        self.J_regressor = torch.tensor(np.concatenate([jr, tips], 1), dtype=dtype)  # (778,21)
        self.hands_comp = torch.tensor(np.asarray(model["hands_components"]), dtype=dtype)
        self.hands_mean = torch.tensor(np.asarray(model["hands_mean"]), dtype=dtype)
        self.weights = torch.tensor(np.asarray(model["weights"]), dtype=dtype)       # (778,16)
        self.parents = PARENTS
        self.coll_mask = collision_mask().to(dtype)


def rodrigues(theta: torch.Tensor) -> torch.Tensor:
    """(N,3) axis-angle -> (N,3,3); epsilon goes *inside* the norm (mano_layer.py:720-728)."""
    n = torch.sqrt(((theta + 1e-8) ** 2).sum(1, keepdim=True))
    axis = theta / n
    half = n * 0.5
    return quat_to_mat(torch.cat([torch.cos(half), torch.sin(half) * axis], 1))


def quat_to_mat(q: torch.Tensor) -> torch.Tensor:
    """(N,4) (w,x,y,z), re-normalised first (mano_layer.py:697-718)."""
    q = q / torch.sqrt((q * q).sum(1, keepdim=True))
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rows = [
        w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
        2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x,
        2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z,
    ]
    return torch.stack(rows, 1).view(-1, 3, 3)


def mano_forward(c: ManoConstants, beta, theta, quat):
    """Restates MANO_SMPL.forward(get_skin=True) (mano_layer.py:573-641).

    Returns verts (B,779,3), joints (B,21,3), Rs (B,15,3,3) in MANO metres."""
    B = beta.shape[0]
    v_shaped = (beta @ c.shapedirs).view(B, 778, 3) + c.v_template                 # :586
    J = torch.einsum("bvc,vj->bjc", v_shaped, c.J_regressor)[:, :16]              # :587-591,:616
    if quat.shape[-1] == 3:
        R0 = rodrigues(quat)                                                       # :595
    else:
        R0 = quat_to_mat(quat)                                                     # :607
    ncomp = theta.shape[-1]
    angles = theta @ c.hands_comp[:ncomp] + c.hands_mean                           # :601
    Rs = rodrigues(angles.reshape(-1, 3)).view(B, 15, 3, 3)
    eye = torch.eye(3, dtype=beta.dtype)
    pose_feature = (Rs - eye).reshape(B, 135)                                      # :611
    v_posed = v_shaped + (pose_feature @ c.posedirs).view(B, 778, 3)               # :613

    # kinematic chain (:730-770): G_i = G_parent(i) . [R_i | J_i - J_parent(i)]
    Rall = torch.cat([R0[:, None], Rs], 1)
    Gr = [Rall[:, 0]]
    Gt = [J[:, 0]]
    for i in range(1, 16):
        p = c.parents[i]
        Gr.append(Gr[p] @ Rall[:, i])
        Gt.append((Gr[p] @ (J[:, i] - J[:, p])[..., None])[..., 0] + Gt[p])
    Gr = torch.stack(Gr, 1)                                                        # (B,16,3,3)
    Gt = torch.stack(Gt, 1)                                                        # (B,16,3)
    At = Gt - (Gr @ J[..., None])[..., 0]                                          # :765-768 rest pose removed
    # linear blend skinning (:619-629)
    Tr = torch.einsum("vj,bjrc->bvrc", c.weights, Gr)
    Tt = torch.einsum("vj,bjr->bvr", c.weights, At)
    verts = (Tr @ v_posed[..., None])[..., 0] + Tt
    joints = torch.einsum("bvc,vj->bjc", verts, c.J_regressor)                     # :630-633
    wrist = verts[:, WRIST_RING].mean(1, keepdim=True)                             # :636
    return torch.cat([verts, wrist], 1), joints, Rs


def get_mano_vertices(c: ManoConstants, quat, pose, shape, cam, global_scale=None):
    """mano_layer.py:643-678: metres -> mm [x global_scale], x cam scale, + cam translation."""
    verts, joints, _ = mano_forward(c, shape, pose, quat)
    s = cam[:, 0].reshape(-1, 1, 1)
    t = cam[:, 1:].reshape(-1, 1, 3)
    joints = joints * 1000
    verts = verts * 1000
    if global_scale is not None:
        joints = joints * global_scale
        verts = verts * global_scale
    return verts * s + t, joints * s + t


def split_params(p):
    """(B,62) -> quat(3), theta(45), beta(10), cam(4) (mano_layer.py:1073-1076);
    a 63-wide vector carries a 4-component quaternion (:988-991)."""
    qd = 4 if p.shape[-1] == 63 else 3
    return p[:, :qd], p[:, qd:qd + 45], p[:, qd + 45:qd + 55], p[:, qd + 55:]


# ----------------------------------------------------------------------------------------
# collision proxy (C1)
# ----------------------------------------------------------------------------------------
def collision_mask() -> torch.Tensor:
    """The constant 66x66 pair table of mano_layer.py:239-269, rebuilt from its rules:
    21 palm spheres (root + 4 per metacarpal) then 3 spheres for each of the 15 finger bones."""
    NP, I = 21, 3
    m = torch.zeros(66, 66)
    m[:NP, NP:] = 1.0                      # palm-palm off, palm-finger on
    m[NP:, :] = 1.0
    for b in range(15):
        root = b // 3 + 1
        rows = slice(NP + I * b, NP + I * b + I)
        if b % 3 == 0:                     # proximal bone of a finger (:252-257)
            m[rows, root * 4] = 0.0
            m[root * 4, rows] = 0.0
            m[rows, NP + I * b: NP + I * b + I + 3] = 0.0
        else:                              # (:259-263)
            lo = NP + I * b - I
            hi = min(NP + I * b + 2 * I + 1, NP + 3 * I * root)
            m[rows, lo:hi] = 0.0
    th = 12 * I                            # thumb root never collides with the palm (:265-269)
    m[NP + th: NP + th + I + 1, :NP] = 0.0
    m[:NP, NP + th: NP + th + I + 1] = 0.0
    return m


def sphere_set(c: ManoConstants, joints, mesh):
    """get_sphere_radius (mano_layer.py:271-317): 66 centres (B,66,3) and radii (B,66)."""
    B = joints.shape[0]
    support = (c.J_regressor > 0).T[None]                                     # (1,21,778)
    d = joints[:, :, None] - mesh[:, None, :778]
    dis = torch.sqrt((d * d).sum(-1) + 1e-8)
    dis = torch.where(support, dis, torch.full_like(dis, 100.0))
    jr = torch.topk(dis, 10, dim=-1, largest=False)[0].mean(-1)
    jr = torch.cat([jr[:, :16], jr[:, [3, 6, 9, 12, 15]] / 1.5], -1)          # :281
    palm_t = torch.tensor([0.2, 0.4, 0.6, 0.8], dtype=joints.dtype)            # :236 linspace(0,1,6)[1:-1]
    fing_t = torch.tensor([0.0, 1.0 / 3, 2.0 / 3], dtype=joints.dtype)         # :231 linspace(0,1,4)[:-1]
    pc = jr[:, [1, 4, 7, 10, 13]]
    pp = torch.clamp(jr[:, 0:1] - 0.05, 0.01, 0.4)                             # :285
    palm_r = ((pc - pp)[..., None] * palm_t + pp[..., None]).reshape(B, -1)
    palm_r = torch.cat([pp, palm_r], 1)
    fc, fp = jr[:, CHILD], jr[:, 1:16]
    fing_r = ((fc - fp)[..., None] * fing_t + fp[..., None]).reshape(B, -1)
    pcc, ppc = joints[:, [1, 4, 7, 10, 13]], joints[:, 0:1]
    palm_c = ((pcc - ppc)[:, :, None] * palm_t[:, None] + ppc[:, :, None]).reshape(B, -1, 3)
    palm_c = torch.cat([ppc, palm_c], 1)
    fcc, fpc = joints[:, CHILD], joints[:, 1:16]
    fing_c = ((fcc - fpc)[:, :, None] * fing_t[:, None] + fpc[:, :, None]).reshape(B, -1, 3)
    return torch.cat([palm_c, fing_c], 1), torch.cat([palm_r, fing_r], 1)


def calculate_coll(c: ManoConstants, joints, mesh):
    """mano_layer.py:373-385.  NB the reference's gate (:383) applies .sum(-1, keepdim=True)
    twice, so the second sum runs over a size-1 axis: the 0.1 threshold gates each sphere
    ROW (b, i), not the whole hand.  The golden vectors pin that behaviour."""
    cen, rad = sphere_set(c, joints, mesh)
    d = cen[:, :, None] - cen[:, None]
    dis = torch.sqrt((d * d).sum(-1) + 1e-8)
    err = torch.clamp(rad[:, :, None] + rad[:, None] - dis, min=0) * c.coll_mask.to(joints.dtype)
    gate = (err.sum(-1, keepdim=True) < 0.1).to(joints.dtype)
    return (err * gate).sum(-1).mean()


# ----------------------------------------------------------------------------------------
# camera / crop helpers (R3, R4, R5 glue) - all torch, float32 like the reference
# ----------------------------------------------------------------------------------------
def points3d_to_img(xyz, intr):
    """mano_layer.py:1318-1324 (note the 1e-8 only on the x divide)."""
    fx, fy, fu, fv = intr
    u = xyz[..., 0] * fx / (xyz[..., 2] + 1e-8) + fu
    v = xyz[..., 1] * fy / xyz[..., 2] + fv
    return torch.stack([u, v, xyz[..., 2]], -1)


def com_to_bounds(com, size, intr):
    """mano_layer.py:1133-1141 -> int32 (xstart, xend, ystart, yend)."""
    fx, fy, _, _ = intr
    xs = torch.floor((com[:, 0] * com[:, 2] / fx - size[:, 0] / 2.) / com[:, 2] * fx + 0.5).int()
    xe = torch.floor((com[:, 0] * com[:, 2] / fx + size[:, 0] / 2.) / com[:, 2] * fx + 0.5).int()
    ys = torch.floor((com[:, 1] * com[:, 2] / fy - size[:, 1] / 2.) / com[:, 2] * fy + 0.5).int()
    ye = torch.floor((com[:, 1] * com[:, 2] / fy + size[:, 1] / 2.) / com[:, 2] * fy + 0.5).int()
    return xs, xe, ys, ye


def offset_to_trans(xs, xe, ys, ye, crop):
    """mano_layer.py:1143-1169 -> M (B,3,3) float32 = off . diag(s,s,1) . trans."""
    wb, hb = xe - xs, ye - ys
    wide = wb > hb
    sz0 = torch.where(wide, torch.full_like(wb, crop), (wb * crop / hb).int())
    sz1 = torch.where(wide, (hb * crop / wb).int(), torch.full_like(wb, crop))
    s = torch.where(wide, crop / wb, crop / hb)
    ox = torch.floor(crop / 2. - sz0 / 2.).int()
    oy = torch.floor(crop / 2. - sz1 / 2.).int()
    M = torch.zeros(xs.shape[0], 3, 3)
    M[:, 0, 0] = s
    M[:, 1, 1] = s
    M[:, 2, 2] = 1
    M[:, 0, 2] = ox - s * xs
    M[:, 1, 2] = oy - s * ys
    return M


def joint_trans(xyz, M, com, cube, intr, crop):
    """mano_layer.py:1301-1309: camera-space points -> crop-normalised uvd."""
    uvd = points3d_to_img(xyz, intr)
    ones = torch.ones_like(uvd[..., :1])
    uv1 = torch.cat([uvd[..., :2], ones], -1)
    uv = torch.einsum("brc,bjc->bjr", M, uv1)[..., :2] / crop * 2 - 1
    d = (uvd[..., 2:] - com[:, None, 2:]) / (cube[:, None, 2:] / 2.0)
    return torch.cat([uv, d], -1)


def normalize_img(z, com_z, cube_z):
    """mano_layer.py:1289-1299 on a raw depth image (0 / -1 = background)."""
    zmin = (com_z - cube_z / 2.).view(-1, 1, 1, 1)
    zmax = (com_z + cube_z / 2.).view(-1, 1, 1, 1)
    z = torch.where((z == -1) | (z == 0), zmax, z)
    z = torch.where(z > zmax, zmax, z)
    z = torch.where(z < zmin, zmin, z)
    return (z - com_z.view(-1, 1, 1, 1)) / (cube_z.view(-1, 1, 1, 1) / 2.)


def literal_sample_maps(M, W, H, S, crop):
    """Which raster pixel does each crop pixel end up reading through the reference's
    resize (S x S -> H x W, mano_layer.py:1233-1242) and warpPerspective
    (mano_layer.py:1244-1260) nearest-neighbour grid_samples?  M is axis-aligned, so the map
    is separable.  Returns int64 (B,crop) column map xi, row map yi (-1 = reads zero padding)
    and bool masks of entries that sit within 1e-3 px of a .5 rounding tie (class T3).
    Evaluated in float64 so ties are detected, not decided by float noise."""
    M = M.double()
    idx = torch.arange(crop, dtype=torch.float64)
    out = []
    for axis, (n_sensor) in enumerate((W, H)):
        s = M[:, axis, axis][:, None]
        t = M[:, axis, 2][:, None]
        u = (idx[None] - t) / s - 0.5                  # sensor pixel coordinate (grid_sample unnormalise)
        r = torch.round(u)                              # half-to-even like nearbyint
        tie = (u - torch.floor(u) - 0.5).abs() < 1e-3
        oob = (r < 0) | (r > n_sensor - 1)
        rc = r.clamp(0, n_sensor - 1)
        src = (2 * rc + 1) * S / (2 * n_sensor) - 0.5  # resize: sensor pixel -> raster pixel
        q = torch.round(src)
        tie2 = (src - torch.floor(src) - 0.5).abs() < 1e-3
        q = q.clamp(0, S - 1).long()
        q[oob] = -1
        out += [q, (tie | tie2) & ~oob]
    return out[0], out[2], out[1], out[3]


# ----------------------------------------------------------------------------------------
# "next" row f1: crop_hand (data/render_loader.py:1209-1227) with uvdImg2xyzImg (:1190-1200),
# uvd_nl2xyz_tensor (:1044-1057), get_trans_points (:1113-1118), pointsImgTo3D (:336-343, flip=1)
# ----------------------------------------------------------------------------------------
def crop_hand_box(joint, center, cube, offsetxy=25.0, offsetz=20.0, hand_thickness=20.0):
    """(B,J,3) normalised teacher joints -> (B,6) [minx,maxx,miny,maxy,minz,maxz] in camera mm."""
    sk = joint * cube[:, None] / 2 + center[:, None]
    lo, hi = sk.min(1)[0], sk.max(1)[0]
    return torch.stack([lo[:, 0] - offsetxy, hi[:, 0] + offsetxy, lo[:, 1] - offsetxy, hi[:, 1] + offsetxy,
                        lo[:, 2] - offsetz - hand_thickness, hi[:, 2] + offsetz], 1)


def crop_hand(img, joint, center, M, cube, intr, offsetxy=25.0, offsetz=20.0, hand_thickness=20.0):
    """img (B,1,R,R) normalised depth -> same with everything outside the skeleton's 3-D box set
    to background (1.0).  M is the axis-aligned crop transform (its inverse is applied in closed form)."""
    B, _, R, _ = img.shape
    fx, fy, px, py = intr
    box = crop_hand_box(joint, center, cube, offsetxy, offsetz, hand_thickness)
    g = 2.0 * torch.arange(R, dtype=img.dtype) / (R - 1.0) - 1.0
    uu = ((g + 1) * (R / 2)).view(1, 1, R)              # column coordinate, crop pixels
    vv = ((g + 1) * (R / 2)).view(1, R, 1)
    d = img[:, 0] * (cube[:, 2] / 2.0).view(B, 1, 1) + center[:, 2].view(B, 1, 1)
    us = (uu - M[:, 0, 2].view(B, 1, 1)) / M[:, 0, 0].view(B, 1, 1)
    vs = (vv - M[:, 1, 2].view(B, 1, 1)) / M[:, 1, 1].view(B, 1, 1)
    x = (us - px) * d / fx
    y = (vs - py) * d / fy
    b = box.view(B, 6, 1, 1)
    mask = (x > b[:, 0]) & (x < b[:, 1]) & (y > b[:, 2]) & (y < b[:, 3]) & (d > b[:, 4]) & (d < b[:, 5])
    return torch.where(mask[:, None], img, torch.ones_like(img)), mask


def seg_pcl(c: ManoConstants, joints, joints_mano, mesh, pcl):
    """mano_layer.py:404-426: 0 = palm, 1..15 = finger bone with the nearest sphere surface."""
    cen, _ = sphere_set(c, joints, mesh)
    _, rad = sphere_set(c, joints_mano, mesh)
    fd = (torch.sqrt(((pcl[:, :, None] - cen[:, None, 21:]) ** 2).sum(-1) + 1e-8) - rad[:, None, 21:]).abs()
    fmin, fid = fd.min(-1)
    pd = (torch.sqrt(((pcl[:, :, None] - cen[:, None, :21]) ** 2).sum(-1) + 1e-8) - rad[:, None, :21]).abs()
    pmin = pd.min(-1)[0]
    bone = (fid.float() / 3).long() + 1
    return torch.where(pmin < fmin, torch.zeros_like(bone), bone)


# ----------------------------------------------------------------------------------------
# "next" row f1 (second half): Img2pcl (data/render_loader.py:1121-1156) and uvdImg2xyzImg
# (:1190-1200) with uvd_nl2xyznl_tensor / uvd_nl2xyz_tensor (:1044-1073).  M is inverted with
# torch.inverse like the reference; flip = loader.flip, img_size = loader.img_size.
# ----------------------------------------------------------------------------------------
def uvd_img_to_xyz(img, center, M, cube, intr, img_size=None, flip=1.0):
    """img (B,1,R,R) normalised depth -> (xyz_img mm, xyz_normal), both (B,3,R,R)."""
    B, _, R, _ = img.shape
    fx, fy, px, py = intr
    img_size = float(R if img_size is None else img_size)
    g = 2.0 * torch.arange(R, dtype=img.dtype) / (R - 1.0) - 1.0
    uu = ((g + 1) * (img_size / 2)).view(1, 1, R).expand(B, R, R)
    vv = ((g + 1) * (img_size / 2)).view(1, R, 1).expand(B, R, R)
    d = img[:, 0] * (cube[:, 2] / 2.0).view(B, 1, 1) + center[:, 2].view(B, 1, 1)
    Mi = torch.inverse(M)
    us = Mi[:, 0, 0].view(B, 1, 1) * uu + Mi[:, 0, 1].view(B, 1, 1) * vv + Mi[:, 0, 2].view(B, 1, 1)
    vs = Mi[:, 1, 0].view(B, 1, 1) * uu + Mi[:, 1, 1].view(B, 1, 1) * vv + Mi[:, 1, 2].view(B, 1, 1)
    x = (us - px) * d / fx
    y = flip * (vs - py) * d / fy
    xyz = torch.stack([x, y, d], 1)
    xyz_n = (xyz - center.view(B, 3, 1, 1)) / (cube.view(B, 3, 1, 1) / 2.0)
    return xyz, xyz_n


def img2pcl_points(img, feature_size, center, M, cube, intr, img_size=None, flip=1.0):
    """The deterministic part of Img2pcl: per hand the (n_b,3) cube-normalised points of the
    foreground cells (value <= 0.99) of the nearest-resized image, in pixel order."""
    B, _, R, _ = img.shape
    img_size = float(R if img_size is None else img_size)       # loader.img_size, NOT feature_size
    img_rs = torch.nn.functional.interpolate(img, (feature_size, feature_size))
    _, xyz_n = uvd_img_to_xyz(img_rs, center, M, cube, intr, img_size=img_size, flip=flip)
    mask = img_rs[:, 0] <= 0.99
    pts = xyz_n.permute(0, 2, 3, 1)
    return [pts[b][mask[b]] for b in range(B)]


def target_from_u16(depth_mm, center, cube, invalid_value=0):
    """loader.normalize_img (data/render_loader.py:738-745) on a (B,R,R) integer-millimetre crop,
    in float32 like the loader's arrays: invalid / zero / far -> far plane, near clamp, (d - cz) / (cube_z / 2)."""
    d = depth_mm.to(torch.float32)
    cz = center[:, 2].view(-1, 1, 1).to(torch.float32)
    hz = (cube[:, 2] / 2.0).view(-1, 1, 1).to(torch.float32)
    far, near = cz + hz, cz - hz
    bad = depth_mm == 0
    if invalid_value:
        bad = bad | (depth_mm == invalid_value)
    d = torch.where(bad, far.expand_as(d), d)
    d = torch.where(d >= far, far.expand_as(d), d)
    d = torch.where(d <= near, near.expand_as(d), d)
    return (d - cz) / hz
