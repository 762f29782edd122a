import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "face-recognition-cpp-tensorrt_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no GPU in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """Build (or reuse) libfr_b200.so; the CPU suite checks that it loads and exports the ABI."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("frb_build", ROOT / "face-recognition-cpp-tensorrt_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not mod.LIB.exists() or os.environ.get("FR_REBUILD"):
        mod.build()
    return mod.LIB
