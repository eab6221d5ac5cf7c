import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200"), os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    g = os.path.join(ROOT, "tests", "golden")
    return json.load(open(os.path.join(g, "kats.json"))), np.load(os.path.join(g, "ref_fixtures.npz"))


@pytest.fixture(scope="session")
def golden_next():
    """Fixtures of the SURVEY 8f blocks (tests/golden/make_golden.py next)."""
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_fixtures_next.npz"))


@pytest.fixture(scope="session")
def orc():
    """The plain-C oracle restatement (built on demand with gcc)."""
    import orc as _orc
    _orc.lib()
    return _orc


@pytest.fixture(scope="session")
def ref():
    """The reference's own classes (oracle/_ref/libgrref.so) when available."""
    import refharness
    if not refharness.available():
        pytest.skip("oracle/_ref/libgrref.so not built (needs /root/reference)")
    refharness.set_fir_impl(1)
    return refharness


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
