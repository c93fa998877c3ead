"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol the header
declares; product code refuses to run without CUDA instead of falling back."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dsf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dsf_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from dsf_b200 import _lib, build

    build.build()
    lib = _lib.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dsf_b200.h but not exported"
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
    assert lib.dsf_version() == 100
    # the workspace is laid out in groups of 8 hands (the TMA row-group view of the blend GEMM operand)
    assert lib.dsf_mano_workspace_floats(24) == 3 * lib.dsf_mano_workspace_floats(8) > 0
    assert lib.dsf_mano_workspace_floats(3) == lib.dsf_mano_workspace_floats(8)


def test_sass_is_sm100a():
    import subprocess

    from dsf_b200 import build

    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(mano_model):
    from dsf_b200.mano_layer import MANO_SMPL
    from dsf_b200.render_loss import m2d_loss

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MANO_SMPL(mano_model, "nyu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m2d_loss(torch.zeros(1, 1, 8, 8), torch.zeros(1, 1, 8, 8))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dsf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "liboracle" not in src, f


def test_row_run_packer_round_trip_cpu():
    """dsf_pack_u16_rows is host code (the loader's side of the row-run transport): its output must reproduce the
    crop wherever the crop is not background, and describe every row by one span."""
    import numpy as np
    import torch

    from dsf_b200.pcl import pack_target_rows

    rng = np.random.RandomState(0)
    B, R = 5, 64
    d = np.zeros((B, R, R), np.uint16)
    center = np.tile(np.array([[0.0, 0.0, 800.0]], np.float32), (B, 1))
    cube = np.full((B, 3), 250.0, np.float32)
    for b in range(B - 1):                                   # blobs of valid depth, holes inside, far-plane pixels
        r0, c0 = rng.randint(5, 30, 2)
        blob = rng.randint(700, 1000, (25, 20)).astype(np.uint16)
        blob[rng.rand(25, 20) < 0.2] = 0
        d[b, r0:r0 + 25, c0:c0 + 20] = blob
    d[0, 10, :] = 0                                          # an empty row inside the blob; hand B-1 is all background
    d[1, 12, 3] = 60000                                      # beyond the far plane: background
    p = pack_target_rows(torch.from_numpy(d), torch.from_numpy(center), torch.from_numpy(cube), pin=False)
    rows, off, pay = p.rows.numpy(), p.hand_offset.numpy(), p.payload.numpy()
    far = (center[:, 2] + cube[:, 2] / np.float32(2)).astype(np.float32)
    fg = (d != 0) & (d.astype(np.float32) < far[:, None, None])
    rebuilt = np.zeros_like(d)
    n = 0
    for b in range(B):
        assert off[b] == n
        for r in range(R):
            c0, ln = rows[b, r]
            cols = np.nonzero(fg[b, r])[0]
            if len(cols) == 0:
                assert ln == 0
                continue
            assert c0 == cols[0] and ln == cols[-1] - cols[0] + 1
            rebuilt[b, r, c0:c0 + ln] = pay[n:n + ln]
            n += ln
    assert off[B] == n and (n == len(pay) or (n == 0 and len(pay) == 1))
    assert np.array_equal(rebuilt[fg], d[fg])
    assert rows[B - 1, :, 1].sum() == 0 and rows[0, 10, 1] == 0
    assert p.nbytes < d.nbytes / 2


def test_entry_points_validate_arguments_before_touching_the_device():
    """Error behaviour of the C ABI (no GPU needed): bad arguments come back as DSF_ERR_BAD_ARG with a message, never
    as a crash or a launch - null handles / buffers, a mis-aligned row-run table, an out-of-range crop size."""
    import ctypes as C

    from dsf_b200 import _lib

    lib = _lib.load_library()
    BAD = -1
    buf = (C.c_float * 64)()
    p = C.cast(buf, C.c_void_p).value
    intr = (C.c_float * 4)(588.0, 587.0, 320.0, 240.0)
    # fused step on the row-run target: null row table
    rc = lib.dsf_fit_step_rows(None, 4, 128, p, p, p, p, p, p, None, None, None, 0, 0.1, 4, None, 0, None, intr,
                               p, None, p, p, p, p, p, p, 0, None)
    assert rc == BAD and b"row-run" in lib.dsf_last_error_string()
    # mis-aligned row table (uint16 pairs are read as 32-bit words)
    rc = lib.dsf_fit_step_rows(None, 4, 128, p, p, p, p, p, p, p + 2, p, p, 0, 0.1, 4, None, 0, None, intr,
                               p, None, p, p, p, p, p, p, 0, None)
    assert rc == BAD and b"aligned" in lib.dsf_last_error_string()
    # invalid marker outside uint16
    rc = lib.dsf_fit_step_rows(None, 4, 128, p, p, p, p, p, p, p, p, p, 70000, 0.1, 4, None, 0, None, intr,
                               p, None, p, p, p, p, p, p, 0, None)
    assert rc == BAD and b"uint16" in lib.dsf_last_error_string()
    # plain fused step: null handle, crop size out of range
    rc = lib.dsf_fit_step(None, 4, 128, p, p, p, p, p, p, p, 0.1, 4, None, 0, None, intr, p, None, p, p, p, p, p, p, 0, None)
    assert rc == BAD
    rc = lib.dsf_view_setup(0, 4, p, p, intr, 640, 480, 1024, None, p, p, p, None, None)
    assert rc == BAD and b"[8,512]" in lib.dsf_last_error_string()
    rc = lib.dsf_target_from_u16_rows(4, 1024, p, p, p, p, p, 0, p, None)
    assert rc == BAD
    # the host-side packer refuses instead of overrunning its payload buffer
    import numpy as np
    d = np.full((1, 8, 8), 700, np.uint16)
    c = np.array([[0, 0, 800.0]], np.float32)
    q = np.full((1, 3), 250.0, np.float32)
    rows, off, pay = np.empty((1, 8, 2), np.uint16), np.empty(2, np.uint32), np.empty(10, np.uint16)
    n = lib.dsf_pack_u16_rows(1, 8, d.ctypes.data, c.ctypes.data, q.ctypes.data, 0, rows.ctypes.data, off.ctypes.data,
                              pay.ctypes.data, pay.size)
    assert n == -1
