"""Random-init MANO model of the real shape (778 vertices / 1538 faces / 16 joints).

MANO_RIGHT.pkl is not redistributable and not present offline, so tests, the
smoke run and the benchmark use a synthetic model with the nine keys the
reference reads (render_model/mano_layer.py:98-149): ``f``, ``v_template``,
``shapedirs``, ``posedirs``, ``J_regressor`` (scipy sparse, 16x778),
``hands_components``, ``hands_mean``, ``kintree_table``, ``weights``.

The rest pose comes from ``assets/hand_topology.npz`` (a hand-shaped mesh with
the MANO vertex/face counts, derived by tools/make_hand_fixture.py), so face
sizes and depth complexity of the rasterised workload are realistic.  Everything
else (blend shapes, PCA pose space, skin weights) is seeded random data with
plausible magnitudes.
"""
from __future__ import annotations

import os
import pickle

import numpy as np

_ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "hand_topology.npz")

# mano_layer.py:147 reads row 0 of kintree_table as the parent list
MANO_PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
WRIST_RING = [121, 214, 215, 279, 239, 234, 92, 38, 122, 118, 117, 119, 120, 108, 79, 78]
TIP_VERTS = [333, 444, 672, 555, 744]


def _adjacency(faces: np.ndarray, n: int) -> np.ndarray:
    a = np.zeros((n, n), dtype=np.float64)
    for i, j in ((0, 1), (1, 2), (2, 0)):
        a[faces[:, i], faces[:, j]] = 1.0
        a[faces[:, j], faces[:, i]] = 1.0
    return a


def make_synthetic_mano(seed: int = 0) -> dict:
    """Return a dict shaped like the unpickled MANO_RIGHT.pkl (plain numpy + scipy sparse)."""
    import scipy.sparse as sp

    rng = np.random.RandomState(seed)
    d = np.load(_ASSET)
    faces = d["faces"].astype(np.int64)
    vj = d["vert_joint"].astype(np.int64)
    loops = d["joint_loop"].astype(np.int64)

    # normalised cube units (1 = 125 mm) -> metres, hand roughly centred on the origin
    v_unit = d["verts"].astype(np.float64)
    j_unit = d["joint_pos"].astype(np.float64)
    centre = 0.5 * (v_unit.min(0) + v_unit.max(0))
    scale = 0.125 * 0.85
    v_template = (v_unit - centre) * scale
    joints = (j_unit - centre) * scale
    nv = v_template.shape[0]

    # --- skin weights: smoothed one-hot of the part labels, top-4, renormalised
    onehot = np.zeros((nv, 16))
    onehot[np.arange(nv), vj] = 1.0
    adj = _adjacency(faces, nv)
    deg = adj.sum(1, keepdims=True)
    w = onehot
    for _ in range(3):
        w = 0.5 * w + 0.5 * (adj @ w) / np.maximum(deg, 1.0)
    kth = np.sort(w, axis=1)[:, -4][:, None]
    w = np.where(w >= kth, w, 0.0)
    w = np.where(w < 0.02, 0.0, w)
    w /= w.sum(1, keepdims=True)

    # --- joint regressor: every joint gets >= 12 support vertices (mano_layer.py:279 sentinel)
    jreg = np.zeros((16, nv))
    for k in range(16):
        dist = np.linalg.norm(v_template - joints[k], axis=1)
        loop = loops[k][loops[k] >= 0]
        if len(loop) >= 3:
            jreg[k, loop] = 0.9 / len(loop)
            extra = [i for i in np.argsort(dist) if i not in set(loop.tolist())][: max(4, 12 - len(loop))]
            jreg[k, extra] = 0.1 / len(extra)
        else:
            near = np.argsort(dist)[:14]
            wk = 1.0 / (dist[near] + 1e-3)
            jreg[k, near] = wk / wk.sum()
    jreg /= jreg.sum(1, keepdims=True)

    # --- shape blend shapes: smooth low-order fields, a few mm per unit beta
    shapedirs = np.zeros((nv, 3, 10))
    p = v_template / np.abs(v_template).max()
    for b in range(10):
        lin = rng.randn(3, 3) * 0.6
        quad = rng.randn(3, 3) * 0.3
        field = p @ lin.T + (p ** 2) @ quad.T + 0.15 * rng.randn(nv, 3)
        shapedirs[:, :, b] = field * 0.003 / (1.0 + 0.3 * b)

    # --- pose blend shapes: corrections localised around the joint whose rotation drives them
    posedirs = np.zeros((nv, 3, 135))
    for k in range(15):
        dist = np.linalg.norm(v_template - joints[k + 1], axis=1)
        local = np.exp(-((dist / 0.02) ** 2))[:, None]
        for e in range(9):
            posedirs[:, :, 9 * k + e] = local * rng.randn(nv, 3) * 0.002 + rng.randn(nv, 3) * 1e-4

    # --- pose PCA space: orthonormal rows, per-angle scale favouring flexion
    q, _ = np.linalg.qr(rng.randn(45, 45))
    ang_scale = np.tile(np.array([0.10, 0.12, 0.45]), 15)
    hands_components = q * ang_scale[None, :]
    hands_mean = np.tile(np.array([0.0, 0.0, 0.25]), 15) + rng.randn(45) * 0.03

    kintree = np.zeros((2, 16), dtype=np.int64)
    kintree[0] = [4294967295] + MANO_PARENTS[1:]
    kintree[1] = np.arange(16)

    return {
        "f": faces.astype(np.uint32),
        "v_template": v_template,
        "shapedirs": shapedirs,
        "posedirs": posedirs,
        "J_regressor": sp.csc_matrix(jreg),
        "hands_components": hands_components,
        "hands_mean": hands_mean,
        "kintree_table": kintree,
        "weights": w,
    }


_J0_CACHE = {}


def _rest_wrist_joint(seed: int = 0) -> np.ndarray:
    if seed not in _J0_CACHE:
        m = make_synthetic_mano(seed)
        _J0_CACHE[seed] = np.asarray(m["J_regressor"].toarray()[0] @ m["v_template"])
    return _J0_CACHE[seed]


def write_mano_pkl(directory: str, seed: int = 0) -> str:
    """Write ``<directory>/MANO_RIGHT.pkl`` (the file name Render.__init__ appends,
    mano_layer.py:931) and return its path."""
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, "MANO_RIGHT.pkl")
    with open(path, "wb") as f:
        pickle.dump(make_synthetic_mano(seed), f, protocol=2)
    return path


def sample_fit_inputs(batch: int, seed: int = 0, cube_mm: float = 250.0) -> dict:
    """Synthetic per-hand inputs with the distributions of SURVEY.md section 8(d).

    Returns float32 numpy arrays: ``params`` (B,62) = [quat3 | theta45 | beta10 | scale | trans3],
    a perturbed copy ``params_target`` used to render the target depth image,
    ``center3d`` (B,3) mm and ``cube`` (B,3) mm.
    """
    rng = np.random.RandomState(seed)
    quat = rng.uniform(-np.pi, np.pi, (batch, 3))
    theta = np.clip(rng.randn(batch, 45), -2, 2)
    beta = rng.randn(batch, 10)
    scale = rng.uniform(0.9, 1.1, (batch, 1))
    # MANO rotates about the wrist joint; translate so the rotated hand stays centred in the
    # cube (as a real crop around the hand's centre of mass would be), plus the +-0.1 jitter
    j0 = _rest_wrist_joint()
    ang = np.linalg.norm(quat, axis=1, keepdims=True)
    ax = quat / np.maximum(ang, 1e-12)
    v = -j0[None]
    rot = (v * np.cos(ang) + np.cross(ax, v) * np.sin(ang)
           + ax * (ax * v).sum(1, keepdims=True) * (1 - np.cos(ang)))
    trans = -(rot + j0[None]) * scale * 8.0 + rng.uniform(-0.1, 0.1, (batch, 3))
    params = np.concatenate([quat, theta, beta, scale, trans], 1)
    tgt = params.copy()
    tgt[:, :48] += rng.randn(batch, 48) * 0.1
    center = np.stack(
        [rng.uniform(-100, 100, batch), rng.uniform(-100, 100, batch), rng.uniform(500, 1200, batch)], 1
    )
    cube = np.repeat(cube_mm * (1.0 + rng.uniform(-0.2, 0.2, (batch, 1))), 3, axis=1)
    return {
        "params": params.astype(np.float32),
        "params_target": tgt.astype(np.float32),
        "center3d": center.astype(np.float32),
        "cube": cube.astype(np.float32),
    }


def quantise_depth_mm(img_norm, center3d, cube):
    """Normalised depth crop (B,R,R) -> what a depth sensor would have delivered for it: integer
    millimetres as uint16, 0 where there is no surface in the cube (background = 1.0).  Used by the
    tests and the bench to make sensor-format targets from rendered images."""
    import torch

    B = img_norm.shape[0]
    img = img_norm.reshape(B, -1)
    d = img * (cube[:, 2:3] / 2.0) + center3d[:, 2:3]
    mm = torch.where(img >= 0.99, torch.zeros_like(d), d.round().clamp(1, 65535))
    return mm.to(torch.int32).to(torch.uint16).reshape(img_norm.shape)
