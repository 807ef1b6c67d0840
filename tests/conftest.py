import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

ART = os.path.join(ROOT, "artifacts", "zkCensus", "dev", "160")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcudart.so.12")
        n = ctypes.c_int(0)
        return cuda.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no GPU visible")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def art_dir():
    if not os.path.exists(os.path.join(ART, "proving_key.zkey")):
        pytest.skip("dev artifacts not generated (run __graft_entry__.build())")
    return ART
