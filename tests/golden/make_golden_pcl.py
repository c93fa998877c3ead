"""Generate tests/golden/pcl_golden.npz by running the reference's OWN loader methods
(data/render_loader.py: Img2pcl :1121-1156, uvdImg2xyzImg :1190-1200) on the crop images already
stored in mano_golden.npz.  Works only where /root/reference exists; the vectors are committed.

    python tests/golden/make_golden_pcl.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
NYU = (588.03, 587.07, 320.0, 240.0)


def main():
    from oracle.ref_import import import_reference_loader_module
    rl = import_reference_loader_module()
    ld = rl.loader.__new__(rl.loader)
    ld.img_size, ld.paras, ld.flip = 128, NYU, 1
    g = np.load(os.path.join(ROOT, "tests", "golden", "mano_golden.npz"))
    img = torch.tensor(g["crop_in"])                  # (B,1,128,128) normalised depth renders
    center, cube, M = torch.tensor(g["center3d"]), torch.tensor(g["cube"]), torch.tensor(g["crop_M"])
    B = img.shape[0]
    # a rotated crop transform for the last hand (rotation augmentation, loader 'rot' mode) so the
    # general 3x3 inverse is exercised, and one empty image
    M = M.clone()
    a = np.deg2rad(25.0)
    rot = torch.tensor([[np.cos(a), -np.sin(a), 64 - 64 * np.cos(a) + 64 * np.sin(a)],
                        [np.sin(a), np.cos(a), 64 - 64 * np.sin(a) - 64 * np.cos(a)],
                        [0, 0, 1]], dtype=torch.float32)
    M[B - 1] = rot @ M[B - 1]
    img = img.clone()
    img[1] = 1.0                                      # empty crop
    # the image is mano_golden.npz[crop_in] with hand 1 blanked; tests rebuild it the same way
    out = {"M": M.numpy(), "center3d": center.numpy(), "cube": cube.numpy()}
    for fs in (128, 64):
        pts, cnt = [], []
        for b in range(B):
            if int((torch.nn.functional.interpolate(img[b:b + 1], (fs, fs)) <= 0.99).sum()) == 0:
                cnt.append(0)
                continue
            p = ld.Img2pcl(img[b:b + 1], fs, center[b:b + 1], M[b:b + 1], cube[b:b + 1], sample_num=0)[0]
            pts.append(p.numpy())
            cnt.append(p.shape[0])
        out[f"pts_{fs}"] = np.concatenate(pts, 0)
        out[f"cnt_{fs}"] = np.array(cnt, np.int32)
    # the sampled mode with more samples than points and with fewer: only set properties are
    # comparable (torch.multinomial's stream is not reproduced), store the reference's own output
    torch.manual_seed(5)
    out["sampled_2048"] = ld.Img2pcl(img[[0, 2]], 128, center[[0, 2]], M[[0, 2]], cube[[0, 2]], 2048).numpy()
    out["sampled_3000"] = ld.Img2pcl(img[0:1], 128, center[0:1], M[0:1], cube[0:1], 3000).numpy()
    out["sampled_empty"] = ld.Img2pcl(img[1:2], 128, center[1:2], M[1:2], cube[1:2], 16).numpy()
    xyz, xyz_n = ld.uvdImg2xyzImg(img, center, M, cube)
    out["xyz_sub"] = xyz.reshape(B, 3, -1)[:, :, ::7].numpy()
    out["xyzn_sub"] = xyz_n.reshape(B, 3, -1)[:, :, ::7].numpy()
    # loader.normalize_img (:738-745) on integer-millimetre crops (the sensor format), two hands, a
    # 32-row slab each; the far plane, zeros, a premax marker and values beyond both clamps all occur
    from dsf_b200.synthetic import quantise_depth_mm
    mm = quantise_depth_mm(torch.tensor(g["crop_in"])[:, 0], center, cube).numpy()[[0, 3], 40:72].copy()
    mm[:, 0, :8] = np.array([0, 1, 300, 65535, 2000, 31000, 5, 0], np.uint16)
    premax = 31000
    norm = []
    for k, b in enumerate((0, 3)):
        norm.append(ld.normalize_img(premax, mm[k].astype(np.float32), center[b].numpy(), cube[b].numpy()))
    out["u16_mm"], out["u16_norm"], out["u16_hands"], out["u16_premax"] = mm, np.stack(norm), np.array([0, 3]), premax
    path = os.path.join(ROOT, "tests", "golden", "pcl_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; counts", out["cnt_128"], out["cnt_64"])


if __name__ == "__main__":
    main()
