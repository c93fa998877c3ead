"""TEST INFRASTRUCTURE ONLY - import the reference's own MANO layer, unmodified.

Works only where /root/reference exists (the build container, not the GPU box).
Used by tests/golden/make_golden.py to produce the committed golden vectors and
by CPU tests that cross-check the oracle restatement against the real thing.

Recipe: render_model/mano_layer.py needs ``np.float``
(removed from numpy >= 1.24, used at mano_layer.py:102,112,...), imports
pytorch3d at module scope (mano_layer.py:896-904, not installed here) and
torchvision RoIAlign (mano_layer.py:33).  We alias ``np.float`` and register
empty stand-in modules for the pytorch3d names so the *rest* of the file - the
MANO_SMPL class and Render's pure-torch helpers - runs as shipped.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "render_model", "mano_layer.py"))


def import_reference_mano_module():
    import numpy as np

    if not hasattr(np, "float"):
        np.float = float  # noqa: NPY001 - the reference predates numpy 1.24
    names = {
        "pytorch3d": [],
        "pytorch3d.renderer": [
            "PerspectiveCameras", "RasterizationSettings", "MeshRasterizer", "Textures",
            "TexturesVertex", "MeshRenderer", "BlendParams", "softmax_rgb_blend",
        ],
        "pytorch3d.structures": ["Pointclouds", "Meshes"],
        "pytorch3d.structures.meshes": ["Meshes"],
        "pytorch3d.loss": ["chamfer_distance"],
        "pytorch3d.ops": ["sample_points_from_meshes"],
    }
    for mod, attrs in names.items():
        if mod not in sys.modules:
            m = types.ModuleType(mod)
            for a in attrs:
                setattr(m, a, type(a, (), {}))
            sys.modules[mod] = m
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            sys.modules["cv2"] = types.ModuleType("cv2")
    try:
        import torchvision.ops  # noqa: F401
    except Exception:
        tv = types.ModuleType("torchvision")
        tvo = types.ModuleType("torchvision.ops")
        tvo.RoIAlign = type("RoIAlign", (), {"__init__": lambda self, *a, **k: None})
        tv.ops = tvo
        sys.modules["torchvision"] = tv
        sys.modules["torchvision.ops"] = tvo
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    return importlib.import_module("render_model.mano_layer")


def make_reference_render(mod, mano_layer, cam_para, image_size, crop_size=(128, 128)):
    """Build a reference ``Render`` object without running its __init__ (which needs
    pytorch3d cameras and CUDA, mano_layer.py:939-952,972,977); only the pure-torch
    helpers (comToBounds, Offset2Trans, affine_grid, warpPerspective, resize,
    normalize_img, JointTrans, points3DToImg) are usable on it."""
    import numpy as np
    import torch

    r = mod.Render.__new__(mod.Render)
    torch.nn.Module.__init__(r)
    r.mano_layer = mano_layer
    r.paras = cam_para
    r.img_size = image_size
    r.crop_size = crop_size
    xx, yy = np.meshgrid(np.arange(crop_size[0]), np.arange(crop_size[0]))
    padd = np.ones([crop_size[0], crop_size[0]])
    r.crop_mesh = torch.from_numpy(np.stack((xx, yy, padd), axis=-1).reshape([1, -1, 3])).float()
    return r


def import_reference_loader_module():
    """data/render_loader.py (for crop_hand / uvdImg2xyzImg): additionally needs matplotlib, PIL,
    tensorboardX, ... at import time only; permissive stand-ins are registered for whatever is missing."""
    import_reference_mano_module()

    class _Dummy:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Dummy()

        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Dummy()

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Dummy

    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "mpl_toolkits",
                 "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d", "PIL", "PIL.Image", "tensorboardX",
                 "prefetch_generator", "trimesh", "aabbtree"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Any(name)
    import importlib

    return importlib.import_module("data.render_loader")
