"""The code paths behind the switches (staged boxes off): every kernel that stages its input through TMA boxes also has a
direct-load path - taken when tensor maps cannot be encoded, and selectable with DS_SRC_BOX / DS_ACC_RING / DS_PYR_BOX /
DS_GAP_FULL = 0 for A/B runs. The switches are read once per process, so the cases run in a child process: on the emulator
here, and (marked gpu) on the device."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["small_mb5", "mb8_crops_bands", "many_frames_one_spot", "bands8_and_row_bands", "affine_f64", "homography_f64"]

CHILD = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from drone_image_stitch_cpp_b200 import _lib
import parity_cases as P
lib = _lib.default_library() if {gpu} else _lib.Library(os.path.join({root!r}, "tests", "emu", "_build", "libdronestitch_emu.so"))
for name in {cases!r}:
    P.CASES[name](lib)
print("FALLBACK_OK")
"""


def _run(gpu):
    env = dict(os.environ, DS_SRC_BOX="0", DS_ACC_RING="0", DS_PYR_BOX="0", DS_GAP_FULL="0")
    out = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT, gpu=gpu, cases=CASES)], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "FALLBACK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_direct_load_paths_emu(emu_lib):
    _run(False)


@pytest.mark.gpu
def test_direct_load_paths_gpu(cuda_lib):
    _run(True)
