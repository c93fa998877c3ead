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
    assert lib.dsf_mano_workspace_floats(3) == 3 * lib.dsf_mano_workspace_floats(1) > 0


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
