import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import kvazzup_b200
        return kvazzup_b200.lib().b200_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly rather than pass on a fallback,
    # so GPU tests are NOT auto-skipped; they are only deselected by `-m "not gpu"`.
    return


@pytest.fixture(scope="session")
def b200():
    import kvazzup_b200
    return kvazzup_b200.lib()


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    return oracle.load()
