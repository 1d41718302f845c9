"""Generates tests/golden/*.npz with the REAL OpenCV (cv2 wheel of this image, 4.13.0) — the library the
reference's hot path runs in (/root/reference/CMakeLists.txt:18; the reference ships no fixtures of its
own). Inputs are stored with the expected outputs so nothing is re-synthesised on the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

from drone_image_stitch_cpp_b200 import synth  # noqa: E402
from oracle import cv_reference as CR  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def plane_case(name, nx, ny, fw, fh, overlap, seed, ws, blend, bands):
    sv = synth.grid_survey(nx, ny, fw, fh, overlap=overlap, seed=seed, work_scale=ws)
    pano, mask, roi = CR.compose_cv2(sv.frames, sv.Ks, sv.Rs, sv.scale, blend, bands)
    roi0, xm, ym, xy, a = CR.maps_cv2((fw, fh), sv.Ks[0], sv.Rs[0], sv.scale)
    corner0, warped0, wmask0 = CR.warp_frame_cv2(sv.frames[0], sv.Ks[0], sv.Rs[0], sv.scale)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), frames=np.stack(sv.frames), Ks=np.stack(sv.Ks), Rs=np.stack(sv.Rs),
                        scale=np.float32(sv.scale), blend=blend, bands=bands, pano=pano, mask=mask, roi=np.array(roi),
                        xy0=xy, a0=a, corner0=np.array(corner0), warped0=warped0, wmask0=wmask0,
                        cv_version=cv2.__version__)
    print(name, pano.shape, roi)


def pyramid_case():
    rng = np.random.default_rng(123)
    out = {}
    for i, (h, w) in enumerate([(37, 53), (64, 64), (2, 9), (48, 130)]):
        a = rng.integers(-300, 300, (h, w, 3)).astype(np.int16)
        f = rng.random((h, w)).astype(np.float32)
        out[f"s16_{i}"] = a
        out[f"f32_{i}"] = f
        out[f"down16_{i}"] = cv2.pyrDown(a)
        out[f"up16_{i}"] = cv2.pyrUp(a)
        out[f"downf_{i}"] = cv2.pyrDown(f)
    np.savez_compressed(os.path.join(HERE, "pyramids.npz"), **out)


def affine_case():
    rng = np.random.default_rng(77)
    src = rng.integers(0, 256, (120, 160, 3)).astype(np.uint8)
    th, s = 0.21, 1.07
    M = np.array([[s * np.cos(th), -s * np.sin(th), 12.3], [s * np.sin(th), s * np.cos(th), -7.9]])
    H = np.vstack([M, [3e-5, -4e-5, 1.0]])
    full = np.full(src.shape[:2], 255, np.uint8)
    np.savez_compressed(
        os.path.join(HERE, "warp_affine_persp.npz"), src=src, M=M, H=H,
        aff=cv2.warpAffine(src, M, (200, 170), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT),
        aff_mask=cv2.warpAffine(full, M, (200, 170), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT),
        per=cv2.warpPerspective(src, H, (200, 170), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT),
        per_mask=cv2.warpPerspective(full, H, (200, 170), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT))


if __name__ == "__main__":
    plane_case("compose_mb5", 3, 2, 200, 150, 0.6, 31, 0.4, "multiband", 5)
    plane_case("compose_mb3", 2, 2, 180, 140, 0.5, 32, 1.0, "multiband", 3)
    plane_case("compose_feather", 2, 2, 200, 150, 0.6, 33, 0.37, "feather", 0)
    pyramid_case()
    affine_case()
