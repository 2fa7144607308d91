import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def _gpu_available():
    try:
        import imscript_b200 as M
        return M.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without CUDA: the gpu-marked tests are skipped (not
    failed), so a CPU-only run stays readable.  On a GPU box nothing is skipped and
    a missing library is an error, never a skip (there is no CPU fallback)."""
    if not any("gpu" in item.keywords for item in items):
        return
    lib_missing = not os.path.exists(os.path.join(ROOT, "imscript_b200", "lib", "libmorsi_cuda.so"))
    if lib_missing or _gpu_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (libmorsi_cuda has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
