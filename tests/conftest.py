import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402
from oracle import pyoracle  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The host-side package over libneo_b200.so (built on demand)."""
    mod = entry.load_package()
    if not os.path.exists(mod.LIBRARY_PATH):
        entry.build()
    return mod


@pytest.fixture(scope="session")
def gpu(pkg):
    """The package with a device selected. GPU tests must fail loudly, not skip, when the CUDA path is unusable."""
    assert pkg.device_count() > 0, "no CUDA device: -m gpu tests need the B200 box (no CPU fallback exists)"
    pkg.set_device(0)
    return pkg


@pytest.fixture(scope="session")
def orc():
    pyoracle.build()
    return pyoracle.oracle()


@pytest.fixture(scope="session")
def ref():
    """oracle/_ref (the unmodified reference compiled in place) or None when the prebuilt library is absent."""
    return pyoracle.ref()


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "neo_ref_golden.npz"))


def rel_l2(got, want):
    got = np.asarray(got).astype(np.complex128 if np.iscomplexobj(got) or np.iscomplexobj(want) else np.float64)
    want = np.asarray(want)
    den = np.linalg.norm(want)
    return float(np.linalg.norm(got - want) / (den if den > 0 else 1.0))


# north_star tolerances: relative L2 vs neo's own fallback on the same inputs
TOL = {"float32": 1e-5, "float64": 1e-12, "complex64": 1e-5, "complex128": 1e-12}
