"""The C-ABI library builds for sm_100a without a GPU, loads, exports every symbol include/dronestitch.h
declares, and fails loudly (no CPU fallback) when no device is present. No compute calls here."""
import ctypes as C
import os
import re

import pytest

from drone_image_stitch_cpp_b200 import _lib as L
from drone_image_stitch_cpp_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "dronestitch.h")).read()
    return sorted(set(re.findall(r"DS_API\s+[\w\s\*]+?\b(ds_\w+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(L.EXPORTS)


def test_cuda_library_builds_and_exports():
    path = build.build_cuda()
    dll = C.CDLL(path)
    for s in declared_symbols():
        assert hasattr(dll, s), s
    dll.ds_version.restype = C.c_char_p
    assert b"sm_100a" in dll.ds_version()


def test_struct_layouts_match_header():
    # sizes the C compiler sees (checked through a tiny probe compiled from the header)
    import subprocess
    import tempfile
    src = '#include "dronestitch.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(ds_transform), sizeof(ds_canvas_desc), sizeof(ds_canvas_info), sizeof(ds_frame_opts));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")])
        out = subprocess.check_output([os.path.join(d, "p")]).decode().split()
    assert [int(v) for v in out] == [C.sizeof(L.ds_transform), C.sizeof(L.ds_canvas_desc), C.sizeof(L.ds_canvas_info), C.sizeof(L.ds_frame_opts)]


def test_geometry_helper_needs_no_device():
    from drone_image_stitch_cpp_b200 import compositor as CP
    import numpy as np
    lib = L.Library(build.build_cuda())
    xf = CP.plane_transform(np.eye(3), np.eye(3), 1.0)
    assert CP.warp_roi(xf, 640, 480, lib) == (0, 0, 640, 480)


def test_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    lib = L.Library(build.build_cuda())
    d = L.ds_canvas_desc()
    d.width, d.height, d.blend_mode, d.num_bands = 100, 100, L.DS_BLEND_MULTIBAND, 5
    h = C.c_void_p()
    rc = lib.dll.ds_create_canvas(C.byref(d), C.byref(h))
    assert rc == L.DS_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.dll.ds_last_error()


def test_error_paths(emu_lib):
    """Argument validation is host logic shared by both builds; exercised on the emulator here."""
    import numpy as np
    from drone_image_stitch_cpp_b200 import compositor as CP
    with pytest.raises(L.DroneStitchError) as e:
        CP.Canvas((0, 0, 0, 10), lib=emu_lib)
    assert e.value.code == L.DS_ERR_BAD_ARG
    with pytest.raises(L.DroneStitchError):
        CP.Canvas((0, 0, 64, 64), blend="feather", sharpness=0.001, lib=emu_lib)   # window too large
    cv = CP.Canvas((0, 0, 64, 64), "multiband", 3, lib=emu_lib)
    img = np.zeros((32, 32, 3), np.uint8)
    with pytest.raises(L.DroneStitchError) as e:       # frame leaves the canvas ROI
        cv.upload(0, img, CP.plane_transform(np.eye(3), np.array([[1, 0, -500], [0, 1, 0], [0, 0, 1]]), 1.0))
    assert e.value.code == L.DS_ERR_BAD_ARG
    with pytest.raises(L.DroneStitchError) as e:       # download before composite
        cv.download()
    assert e.value.code == L.DS_ERR_STATE
    with pytest.raises(L.DroneStitchError):            # band edges must be multiples of 2^bands
        CP.Canvas((0, 0, 64, 64), "multiband", 3, band=(3, 40), lib=emu_lib)
    cv.upload(0, img, CP.plane_transform(np.eye(3), np.eye(3), 1.0))
    cv.composite()
    with pytest.raises(L.DroneStitchError):            # tile outside the canvas
        cv.download(x=60, y=0, w=10, h=10)
    cv.close()
