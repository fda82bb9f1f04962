import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_shadows():
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "shadow_golden.npz"))
    cases = {}
    for k in ("test1", "test2", "test3", "test4"):
        cases[k] = dict(bhspin=float(g[k + "__bhspin"]), inclination=float(g[k + "__inclination"]),
                        angles=g[k + "__angles"], radii=g[k + "__radii"])
    return cases


@pytest.fixture(scope="session")
def built():
    """Make sure the native artefacts exist (built in-tree; they travel to the GPU box)."""
    import __graft_entry__ as ge
    from mahakala_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH) or not os.path.exists(os.path.join(ROOT, "oracle", "libmk_oracle.so")):
        ge.build()
    return True
