import ctypes
import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from spectrseqtools_b200 import _frame  # noqa: E402

_frame.install_polars_shim()  # only when real polars is absent


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) and the built libsst_b200.so")


def _gpu_count() -> int:
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
            n = ctypes.c_int(0)
            if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0:
                return n.value
            return 0
        except OSError:
            continue
    return 0


HAVE_GPU = _gpu_count() > 0


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Make sure the CUDA library and the C oracle exist (both compile without a GPU)."""
    import __graft_entry__ as g

    g.build()
