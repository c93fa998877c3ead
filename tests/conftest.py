import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "mano_golden.npz"))


@pytest.fixture(scope="session")
def pcl_golden(golden):
    """Img2pcl / uvdImg2xyzImg vectors (tests/golden/make_golden_pcl.py); the input image is the
    crop_in of mano_golden.npz with hand 1 blanked, rebuilt here the way the generator does."""
    import numpy as np

    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "pcl_golden.npz")))
    img = golden["crop_in"].copy()
    img[1] = 1.0
    g["img"] = img
    return g


@pytest.fixture(scope="session")
def mano_model():
    from dsf_b200.synthetic import make_synthetic_mano

    return make_synthetic_mano(0)
