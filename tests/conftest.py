import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle_py import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref).  Built here from /root/reference; travels prebuilt."""
    from oracle.oracle_py import RefHarness
    try:
        return RefHarness()
    except (FileNotFoundError, OSError) as e:  # pragma: no cover
        pytest.skip(f"oracle/_ref not available: {e}")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu but no CUDA device is visible")
    return 0
