import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def emu_lib():
    """TEST-ONLY CPU emulation of the kernel bodies (tests/emu). Never loaded by the product."""
    from drone_image_stitch_cpp_b200 import _lib
    emu_dir = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["make", "-C", emu_dir, "-s"])
    return _lib.Library(os.path.join(emu_dir, "_build", "libdronestitch_emu.so"))


@pytest.fixture(scope="session")
def cuda_lib():
    from drone_image_stitch_cpp_b200 import _lib, build
    build.build_cuda()
    return _lib.default_library()


@pytest.fixture(scope="session")
def small_survey():
    from drone_image_stitch_cpp_b200 import synth
    return synth.grid_survey(3, 2, 400, 300, overlap=0.6, seed=1, work_scale=0.37)
