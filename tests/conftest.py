"""pytest configuration: the `gpu` marker and shared fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle import pyoracle
    pyoracle.build()
    if not pyoracle.have_reference():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return pyoracle.Reference()


@pytest.fixture(params=["auto", "bitslice", "word"])
def matcher(request, monkeypatch):
    """Which K2 serves the scan: the engine's own choice, the bit-sliced kernel forced
    (size thresholds lifted) or the word-parallel kernels forced.  Read by sqbEngineNew."""
    if request.param == "auto":
        monkeypatch.delenv("SEEQ_B200_MATCHER", raising=False)
    else:
        monkeypatch.setenv("SEEQ_B200_MATCHER", request.param)
    return request.param
