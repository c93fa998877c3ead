#!/usr/bin/env python
"""Derive the hand-like template fixture used by make_synthetic_mano().

Reads the 15 watertight part meshes the reference ships under
/root/reference/data/waterObjMap (produced by data/render_loader.py:4825-4953),
merges them by position, drops the cap-centre faces (cap counts per part from
render_loader.py:4866-4871) and recovers the 779-vertex / 1554-face mesh the
reference rasterises (mano_layer.py:102-106).  The vertex with the 16-face fan
is the wrist-cap centre (mano_layer.py:636); removing it leaves the 778/1538
MANO topology.  Vertices are then renumbered so the hard-coded wrist ring
(mano_layer.py:103-105) and fingertip ids (mano_layer.py:125-129) land on the
right geometry.

Run here (the reference tree is not on the GPU box):
    python tools/make_hand_fixture.py
Writes dsf_b200/assets/hand_topology.npz (derived data only: positions, faces,
part labels, joint anchors).
"""
import os
import sys
import numpy as np

REF = "/root/reference/data/waterObjMap"
CAPS = [5, 2, 2, 1, 2, 2, 1, 2, 2, 1, 2, 2, 1, 2, 1]
PART_PARENT = [0, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13]   # eval_coll.py:615
RING_IDS = [121, 214, 215, 279, 239, 234, 92, 38, 122, 118, 117, 119, 120, 108, 79, 78]
TIP_IDS = [333, 444, 672, 555, 744]
DISTAL_PARTS = [3, 6, 9, 12, 14]


def read_obj(path):
    V, F = [], []
    for line in open(path):
        s = line.split()
        if not s:
            continue
        if s[0] == "v":
            V.append([float(x) for x in s[1:4]])
        elif s[0] == "f":
            F.append([int(x) - 1 for x in s[1:4]])
    return np.array(V), np.array(F)


def main(out):
    key2id, verts, faces, face_part = {}, [], [], []
    vert_parts = {}
    caps = []            # (part, centre, loop vertex ids)
    for p in range(15):
        V, F = read_obj(os.path.join(REF, "part%d.obj" % p))
        ncap = CAPS[p]
        cap_local = list(range(len(V) - ncap, len(V)))
        ids = []
        for i, v in enumerate(V):
            if i in cap_local:
                ids.append(-1)
                continue
            k = tuple(np.round(v, 6))
            if k not in key2id:
                key2id[k] = len(verts)
                verts.append(v)
            ids.append(key2id[k])
            vert_parts.setdefault(key2id[k], set()).add(p)
        for c in cap_local:
            loop = set()
            for f in F:
                if c in f:
                    loop.update(ids[i] for i in f if i != c)
            caps.append((p, V[c], sorted(loop)))
        for f in F:
            if any(i in cap_local for i in f):
                continue
            faces.append([ids[i] for i in f])
            face_part.append(p)
    verts = np.array(verts)
    faces = np.array(faces)
    face_part = np.array(face_part)
    assert verts.shape == (779, 3) and faces.shape == (1554, 3), (verts.shape, faces.shape)

    # wrist-cap centre = the vertex carrying the 16-face fan
    cnt = np.bincount(faces.reshape(-1), minlength=779)
    fan_faces = None
    for cand in np.argsort(-cnt):
        ff = np.where((faces == cand).any(1))[0]
        if len(ff) == 16:
            wrist = int(cand)
            fan_faces = ff
            break
    assert fan_faces is not None
    # order the ring by walking the fan: face (a, b, wrist) rotated so wrist is last
    nxt = {}
    for fi in fan_faces:
        f = list(faces[fi])
        k = f.index(wrist)
        a, b = f[(k + 1) % 3], f[(k + 2) % 3]
        nxt[a] = b
    ring = [next(iter(nxt))]
    while len(ring) < 16:
        ring.append(nxt[ring[-1]])
    assert nxt[ring[-1]] == ring[0] and len(set(ring)) == 16

    keep_faces = np.setdiff1d(np.arange(1554), fan_faces)
    faces = faces[keep_faces]
    face_part = face_part[keep_faces]

    # vertex part label: a boundary vertex takes the most distal (largest) part id
    vpart = np.array([max(vert_parts[i]) for i in range(779)])

    # joint anchors: joint k (1..14) = centre of the cap shared by part k and its parent
    joint_pos = np.zeros((16, 3))
    joint_loop = [None] * 16
    for k in range(1, 15):
        mine = [c for c in caps if c[0] == k]
        par = [c for c in caps if c[0] == PART_PARENT[k]]
        best = None
        for a in mine:
            for b in par:
                d = np.linalg.norm(a[1] - b[1])
                if best is None or d < best[0]:
                    best = (d, a)
        assert best[0] < 1e-4, (k, best[0])
        joint_pos[k] = best[1][1]
        joint_loop[k] = best[1][2]

    # fingertip = vertex of the distal part farthest from that part's proximal cap
    tips = []
    for p in DISTAL_PARTS:
        members = [i for i in range(779) if p in vert_parts[i] and i != wrist]
        d = np.linalg.norm(verts[members] - joint_pos[p], axis=1)
        tips.append(members[int(np.argmax(d))])
    # joint 15 splits the last thumb part 45 % of the way to the tip
    joint_pos[15] = joint_pos[14] + 0.45 * (verts[tips[4]] - joint_pos[14])
    axis = verts[tips[4]] - joint_pos[14]
    t = (verts - joint_pos[14]) @ axis / (axis @ axis)
    vpart = np.where((vpart == 14) & (t > 0.45), 15, vpart)
    ring_c = verts[ring].mean(0)
    palm_c = verts[[i for i in range(779) if vert_parts[i] == {0}]].mean(0)
    joint_pos[0] = ring_c + 0.15 * (palm_c - ring_c)

    # renumber: drop the wrist centre, pin ring and tips to the reference's ids
    perm = -np.ones(779, dtype=np.int64)          # old id -> new id
    used = set()
    for o, n in zip(ring, RING_IDS):
        perm[o] = n
        used.add(n)
    for o, n in zip(tips, TIP_IDS):
        assert perm[o] < 0
        perm[o] = n
        used.add(n)
    free = [n for n in range(778) if n not in used]
    it = iter(free)
    for o in range(779):
        if o == wrist or perm[o] >= 0:
            continue
        perm[o] = next(it)
    inv = np.zeros(778, dtype=np.int64)
    for o in range(779):
        if o != wrist:
            inv[perm[o]] = o
    new_verts = verts[inv]
    new_faces = perm[faces]
    assert new_faces.min() >= 0 and new_faces.max() == 777
    new_vpart = vpart[inv]
    loops = np.full((16, 12), -1, dtype=np.int16)
    for k in range(1, 15):
        lp = [perm[i] for i in joint_loop[k]]
        loops[k, :len(lp)] = lp
    np.savez_compressed(
        out,
        verts=new_verts.astype(np.float32),          # normalised units (1 = 125 mm)
        faces=new_faces.astype(np.int16),
        face_part=face_part.astype(np.int8),
        vert_joint=new_vpart.astype(np.int8),         # 0..15
        joint_pos=joint_pos.astype(np.float32),
        joint_loop=loops,
    )
    print("wrote", out, new_verts.shape, new_faces.shape, "tips", [perm[t] for t in tips])


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    main(os.path.join(here, "..", "dsf_b200", "assets", "hand_topology.npz"))
