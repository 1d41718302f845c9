"""The C++ host side (include/dronestitch.hpp: ds::composePanorama, ds::Blender) compiled with g++ and run as the
reference's call site would run it; its panorama must be the oracle's. CPU: linked against the tests-only emulator
build of the library; marked gpu: against libdronestitch_cuda.so."""
import os
import struct
import subprocess

import numpy as np
import pytest

from drone_image_stitch_cpp_b200 import _lib as L
from drone_image_stitch_cpp_b200 import synth
from oracle import ds_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")


def build_driver(lib_path, tag):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, f"host_compose_{tag}")
    src = os.path.join(ROOT, "tests", "cpp", "host_compose.cpp")
    deps = [src, os.path.join(ROOT, "include", "dronestitch.hpp"), os.path.join(ROOT, "include", "dronestitch.h"), lib_path]
    if os.path.exists(exe) and os.path.getmtime(exe) >= max(os.path.getmtime(d) for d in deps):
        return exe
    d, name = os.path.dirname(lib_path), os.path.basename(lib_path)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L", d, "-l:" + name, "-Wl,-rpath," + d])
    return exe


def run_case(exe, tmp_path, blend, bands, work_scale):
    sv = synth.grid_survey(2, 2, 260, 200, overlap=0.55, seed=61, work_scale=0.5)
    # cameras as cv::Stitcher holds them at registration scale: the driver rescales them by 1 / work_scale
    case = os.path.join(tmp_path, "case.bin")
    out = os.path.join(tmp_path, "out.bin")
    with open(case, "wb") as f:
        f.write(b"DSC1")
        f.write(struct.pack("<iiii", len(sv.frames), bands, 1 if blend == "feather" else 0, 1))
        f.write(struct.pack("<f", float(np.float32(sv.scale) * np.float32(work_scale))))
        f.write(struct.pack("<d", work_scale))
        for img, K, R in zip(sv.frames, sv.Ks, sv.Rs):
            K = np.asarray(K, np.float32)
            focal, ppx, ppy = float(K[0, 0]) * work_scale, float(K[0, 2]) * work_scale, float(K[1, 2]) * work_scale
            aspect = float(K[1, 1]) / float(K[0, 0])
            f.write(struct.pack("<ii", img.shape[1], img.shape[0]))
            f.write(struct.pack("<dddd", focal, aspect, ppx, ppy))
            f.write(np.ascontiguousarray(R, np.float32).tobytes())
            f.write(np.ascontiguousarray(img).tobytes())
    r = subprocess.run([exe, case, out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    x, y, w, h = struct.unpack("<iiii", raw[:16])
    pano = np.frombuffer(raw, np.uint8, w * h * 3, 16).reshape(h, w, 3)
    mask = np.frombuffer(raw, np.uint8, w * h, 16 + w * h * 3).reshape(h, w)
    # what the driver must have derived: K scaled by 1 / work_scale (back to sv.Ks), scale = sv.scale
    Ks = []
    for K in sv.Ks:
        K = np.asarray(K, np.float32)
        a = 1.0 / work_scale
        focal, ppx, ppy = float(K[0, 0]) * work_scale * a, float(K[0, 2]) * work_scale * a, float(K[1, 2]) * work_scale * a
        aspect = float(K[1, 1]) / float(K[0, 0])
        Ks.append(np.array([[focal, 0, ppx], [0, focal * aspect, ppy], [0, 0, 1]], np.float64).astype(np.float32))
    scale = np.float32(np.float64(np.float32(sv.scale) * np.float32(work_scale)) * (1.0 / work_scale))
    ref, refmask, roi = O.compose_port(sv.frames, Ks, sv.Rs, scale, blend, bands)
    assert (x, y, w, h) == tuple(roi)
    assert np.array_equal(mask, refmask)
    assert np.array_equal(pano, ref)


def run_global_stage(exe, tmp_path):
    """stitchInterStripsCustom's compose half through the C++ header, against the oracle chain."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from parity_cases import global_stage_specs
    from helpers import assert_blend_parity
    specs = global_stage_specs(seed=77, n=3, fw=380, fh=270)
    rng = np.random.default_rng(77)
    # global transforms before the canvas shift: what estimatePairAffine chains produce (arbitrary origin)
    Hs = []
    for s in specs:
        Hm = np.eye(3)
        Hm[:2] = s["M"]
        Hm[0, 2] += s["corner"][0] - 1234.0
        Hm[1, 2] += s["corner"][1] + 321.0
        Hs.append(Hm)
    gains = [(1.0, 1.0, 1.0)] + [tuple(float(v) for v in rng.uniform(0.9, 1.1, 3)) for _ in specs[1:]]
    case, out = os.path.join(tmp_path, "g.bin"), os.path.join(tmp_path, "g.out")
    with open(case, "wb") as f:
        f.write(b"DSG1")
        f.write(struct.pack("<ii", len(specs), 5))
        for s, Hm, g in zip(specs, Hs, gains):
            img = s["img"]
            f.write(struct.pack("<ii", img.shape[1], img.shape[0]))
            f.write(np.ascontiguousarray(Hm, np.float64).tobytes())
            f.write(np.asarray(g, np.float32).tobytes())
            f.write(np.ascontiguousarray(img).tobytes())
            m = s["seam_lowres"]
            f.write(struct.pack("<ii", m.shape[1], m.shape[0]))
            f.write(np.ascontiguousarray(m).tobytes())
    r = subprocess.run([exe, case, out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    # the oracle side, restating :439-486 and :643-666
    def tbr(w, h, Hm):
        pts = Hm @ np.array([[0, w, w, 0], [0, 0, h, h], [1, 1, 1, 1]], np.float64)
        x, y = int(np.floor(pts[0].min())), int(np.floor(pts[1].min()))
        return x, y, max(1, int(np.ceil(pts[0].max())) - x), max(1, int(np.ceil(pts[1].max())) - y)
    pre = [tbr(s["img"].shape[1], s["img"].shape[0], Hm) for s, Hm in zip(specs, Hs)]
    min_x, min_y = min(p[0] for p in pre), min(p[1] for p in pre)
    raw = open(out, "rb").read()
    off = 0
    warped, masks, corners = [], [], []
    for s, Hm, g in zip(specs, Hs, gains):
        S = Hm.copy()
        S[0, 2] += float(-min_x)
        S[1, 2] += float(-min_y)
        x, y, bw, bh = tbr(s["img"].shape[1], s["img"].shape[0], S)
        M = S[:2].copy()
        M[0, 2] -= float(x)
        M[1, 2] -= float(y)
        content = O.content_mask(s["img"], M, bw, bh)
        got = np.frombuffer(raw, np.uint8, bw * bh, off).reshape(bh, bw)
        off += bw * bh
        assert np.array_equal(got, content)
        wimg = O.remap_bilinear(s["img"], *O.affine_tables(M, bw, bh), "constant")
        wimg = np.clip(np.rint(wimg.astype(np.float32) * np.asarray(g, np.float32)[None, None, :]), 0, 255).astype(np.uint8)
        seam = O.threshold_gt1(O.resize_nearest(s["seam_lowres"], bw, bh))
        warped.append(wimg)
        masks.append(O.soft_blend_mask(seam, content, 10.0))
        corners.append((x, y))
    roi = O.result_roi(corners, [(w.shape[1], w.shape[0]) for w in warped])
    # stitch_global.cpp:632-635 on canvas_w / canvas_h of :455-456 (configured bands: 5)
    canvas_w, canvas_h = max(p[0] + p[2] for p in pre) - min_x, max(p[1] + p[3] for p in pre) - min_y
    final_bands = max(max(5, 5), min(12, int(np.ceil(np.log2(float(max(canvas_w, canvas_h))))) - 1))
    assert final_bands > 5, "the case should exercise the automatic band count"
    bl = O.MultiBand(roi, final_bands)
    for wimg, m, c in zip(warped, masks, corners):
        bl.feed(wimg.astype(np.int16), m, c)
    ref16, refmask = bl.blend()
    x, y, w, h = struct.unpack("<iiii", raw[off:off + 16])
    assert (x, y, w, h) == tuple(roi)
    pano = np.frombuffer(raw, np.uint8, w * h * 3, off + 16).reshape(h, w, 3)
    mask = np.frombuffer(raw, np.uint8, w * h, off + 16 + w * h * 3).reshape(h, w)
    assert np.array_equal(mask, refmask)
    assert_blend_parity(pano, O.s16_to_u8(ref16))
    # the crop: rectangle as the oracle's autoCropBlackBorder picks it (or refused), and pano(rect) downloaded on its own
    off2 = off + 16 + w * h * 4
    kx, ky, kw, kh, decided = struct.unpack("<iiiii", raw[off2:off2 + 20])
    if decided:
        assert (kx, ky, kw, kh) == O.auto_crop_rect(pano)
    cropped = np.frombuffer(raw, np.uint8, kw * kh * 3, off2 + 20).reshape(kh, kw, 3)
    assert np.array_equal(cropped, pano[ky:ky + kh, kx:kx + kw])


def test_cpp_global_stage_emu(emu_lib, tmp_path):
    run_global_stage(build_driver(emu_lib.path, "emu"), str(tmp_path))


@pytest.mark.gpu
def test_cpp_global_stage_gpu(cuda_lib, tmp_path):
    run_global_stage(build_driver(cuda_lib.path, "cuda"), str(tmp_path))


@pytest.mark.parametrize("blend,bands,work_scale", [("multiband", 4, 1.0), ("multiband", 5, 0.5), ("feather", 0, 1.0)])
def test_cpp_host_emu(emu_lib, tmp_path, blend, bands, work_scale):
    run_case(build_driver(emu_lib.path, "emu"), str(tmp_path), blend, bands, work_scale)


@pytest.mark.gpu
@pytest.mark.parametrize("blend,bands,work_scale", [("multiband", 5, 0.5), ("feather", 0, 1.0)])
def test_cpp_host_gpu(cuda_lib, tmp_path, blend, bands, work_scale):
    run_case(build_driver(cuda_lib.path, "cuda"), str(tmp_path), blend, bands, work_scale)
