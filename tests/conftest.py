import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v2_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _ensure_built():
    """Build the CPU-side artefacts the suite needs (cross-compiles CUDA without a GPU)."""
    import subprocess
    lib = os.path.join(ROOT, "lightmetrica-v2_b200", "lib", "liblmb200.so")
    if not os.path.exists(lib):
        subprocess.check_call(["bash", os.path.join(ROOT, "build.sh")])
    orc = os.path.join(ROOT, "oracle", "liblmoracle.so")
    if not os.path.exists(orc):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "liblmoracle.so")])
    if os.path.isdir("/root/reference/src") and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liblightmetrica.so")):
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "ref", "build_ref.sh")])


_ensure_built()

# one libnccl.so.2 per process: the multi-GPU entry points must pick the copy torch will load later (see lmb200py/capi.py)
from lmb200py import capi as _capi      # noqa: E402
_capi._point_at_torch_nccl()


@pytest.fixture(scope="session")
def have_gpu():
    from lmb200py import capi
    return capi.lib().lmb200_device_count() > 0
