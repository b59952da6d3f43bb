import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle
    oracle.build()
    oracle.lib().oracle_set_variants(0, 0)
    oracle.lib().oracle_set_experiment.argtypes = None
    oracle.lib().oracle_set_experiment(0, 0)
    return oracle


@pytest.fixture(scope="session")
def ctx():
    """Device context through the C ABI. Fails loudly when the CUDA library or the GPU is missing."""
    from tray_b200 import ray
    return ray.default_context()
