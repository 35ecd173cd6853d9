import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def lib_path():
    """Path of the built C-ABI library; builds it when nvcc is available and it is missing."""
    from gan_aug_pfa_b200 import _lib
    if not _lib.LIB_PATH.exists():
        import importlib.util
        spec = importlib.util.spec_from_file_location("gap_build", ROOT / "gan-aug-pfa_b200" / "build.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return _lib.LIB_PATH
