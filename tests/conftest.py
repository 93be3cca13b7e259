import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def ctx():
    """The CUDA context. GPU tests FAIL (not skip) when the library or the GPU is missing: a silent
    fallback would void every parity claim."""
    import halo2_snark_aggregator_b200 as h2

    c = h2.Context(0)
    yield c
    c.close()
