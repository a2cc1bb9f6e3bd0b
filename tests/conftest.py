import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
_REF = os.path.join(ROOT, "oracle", "_ref")        # the real pyabpoa / conk when `make -C oracle ref` could build them
if os.path.isdir(_REF) and _REF not in sys.path:
    sys.path.insert(1, _REF)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", params=["auto", "grp", "lane"])
def gpu(request):
    """The CUDA handle.  No skip: on a GPU box a missing library/device must FAIL the test.
    Every GPU test runs three times: "auto" (the group kernel for the bulk of short subreads in batches of >= 12 000
    reads, else the warp-per-read kernel), "grp" (the group kernel -- 8 lanes per read for the DP, one thread per read
    for the graph phases -- for everything it covers, the warp kernel takes what it declines) and "lane" (the
    thread-per-read lane kernel of round 1, same fallback)."""
    from c3poa_b200.api import GpuConsensus
    h = GpuConsensus(0, poa_mode=request.param)
    h.poa_mode = request.param
    yield h
    h.close()


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle
