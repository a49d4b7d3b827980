import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "slow: long-running parity runs (minutes); still part of -m gpu")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return {k: np.load(os.path.join(d, k + ".npz")) for k in ("test1_fast", "pyrdown_cv2", "circle_cv2", "undistort_cv2", "clahe_cv2")}


@pytest.fixture(scope="session")
def built():
    """CUDA library + oracle are built (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def scenario(built):
    import helpers as H
    return H.make_scenario(20260101)


@pytest.fixture(scope="session")
def ctx(built):
    """One context for the GPU tests (kinect geometry, 5 levels). Fails loudly when no CUDA device is present."""
    from dsdtm_b200 import capi, synth as S
    c = capi.Context(S.KINECT, levels=5, cell_size=15, max_feats=320, max_patches=320, max_frames=12, max_batch=8)
    yield c
    c.close()
