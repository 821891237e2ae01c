import json
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))
GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def fh(s):
    """float.fromhex with pass-through for the markers used in the fixtures."""
    return float.fromhex(s)


@pytest.fixture(scope="session")
def golden_math():
    return json.loads((GOLDEN / "ref_math_utils.json").read_text())


@pytest.fixture(scope="session")
def golden_grid():
    return json.loads((GOLDEN / "ref_simplegrid.json").read_text())


@pytest.fixture(scope="session")
def golden_pipeline():
    return json.loads((GOLDEN / "ref_pipeline.json").read_text())
