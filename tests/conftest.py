import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _cuda_devices():
    try:
        from scs_python_b200 import _scs_b200 as B
        return B.lib.scs_b200_device_count()
    except Exception:
        return 0


@pytest.fixture(scope="session")
def gpu():
    n = _cuda_devices()
    if n <= 0:
        pytest.fail("gpu-marked test started without a CUDA device: the B200 backend has no CPU fallback")
    return n


@pytest.fixture(scope="session")
def ref_scs():
    """The compiled reference python package (oracle/_ref), if it travelled to this box."""
    p = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(p, "scs", "__init__.py")):
        pytest.skip("oracle/_ref not built on this box")
    sys.path.insert(0, p)
    try:
        import scs
    except Exception as e:  # pragma: no cover
        pytest.skip("reference package not importable: %r" % (e,))
    return scs
