import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")
    # a fresh checkout has no libcpimh.so (built artefacts are git-ignored): build it once (nvcc cross-compiles without a GPU)
    lib = os.path.join(ROOT, "unbiased_pimh_b200", "libcpimh.so")
    if not os.path.exists(lib):
        import shutil
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            from unbiased_pimh_b200 import build as b
            b.build_library()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle
