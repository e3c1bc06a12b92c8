import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count() -> int:
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu on the B200 box.  Without a device they cannot run;
    # they are skipped (never silently passed through a CPU path — there is none).
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    """The product package (fluid-rs_b200/) with its CUDA library built."""
    import __graft_entry__ as ge
    ge.build()
    import fluidpkg
    return fluidpkg.load()


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def scenes():
    import fluidpkg
    return fluidpkg.load().scenes


def by_id(records, ids):
    o = np.argsort(ids, kind="stable")
    return records[o], ids[o]
