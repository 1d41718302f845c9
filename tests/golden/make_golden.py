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


def global_stage_case():
    """stitchInterStripsCustom's compose half through cv2 itself (src/stitch_global.cpp:470-486, :643-666 and
    src/stitch_common.cpp:4-27): warpAffine, buildWarpedContentMask, seam mask NEAREST resize + threshold,
    buildSoftBlendMask, applyChannelGainInPlace, MultiBandBlender, ->8U, autoCropBlackBorder's rectangle."""
    rng = np.random.default_rng(55)
    fw, fh, n, bands = 300, 200, 3, 4
    ortho = synth.orthophoto(fh * 2 + 100, fw * 2 + 200, 55).numpy()
    strips, Ms, corners, sizes, seams, gains = [], [], [], [], [], []
    for i in range(n):
        th, sc = rng.uniform(-0.12, 0.12), rng.uniform(0.93, 1.07)
        Hm = np.array([[sc * np.cos(th), -sc * np.sin(th), 30 + i * fw * 0.5], [sc * np.sin(th), sc * np.cos(th), 20 + (i % 2) * fh * 0.35], [0, 0, 1.0]])
        pts = Hm @ np.array([[0, fw, fw, 0], [0, 0, fh, fh], [1, 1, 1, 1]], np.float64)
        x0, y0 = int(np.floor(pts[0].min())), int(np.floor(pts[1].min()))
        bw, bh = max(1, int(np.ceil(pts[0].max())) - x0), max(1, int(np.ceil(pts[1].max())) - y0)
        M = Hm[:2].copy(); M[0, 2] -= x0; M[1, 2] -= y0
        img = np.ascontiguousarray(ortho[15 * i:15 * i + fh, 25 * i:25 * i + fw]).copy()
        yy, xx = np.mgrid[0:fh, 0:fw]
        img[(yy < 0.1 * xx - 6 * i) | (yy > fh - 10 + 0.03 * xx)] = 0
        img[60:70, 100:140] = rng.integers(0, 5, (10, 40, 3), dtype=np.uint8)
        low = np.full((max(4, bh // 6), max(4, bw // 6)), 255, np.uint8)
        ly, lx = np.mgrid[0:low.shape[0], 0:low.shape[1]]
        low[(lx > 0.62 * low.shape[1] + 0.15 * ly) if i % 2 == 0 else (lx < 0.3 * low.shape[1] - 0.1 * ly)] = 0
        low[1, 2] = 1
        strips.append(img); Ms.append(M); corners.append((x0, y0)); sizes.append((bw, bh)); seams.append(low)
        gains.append((1.0, 1.0, 1.0) if i == 0 else tuple(float(v) for v in rng.uniform(0.85, 1.2, 3)))
    x_min, y_min = min(c[0] for c in corners), min(c[1] for c in corners)
    x_max, y_max = max(c[0] + s_[0] for c, s_ in zip(corners, sizes)), max(c[1] + s_[1] for c, s_ in zip(corners, sizes))
    blender = cv2.detail_MultiBandBlender(0, bands)
    blender.prepare((x_min, y_min, x_max - x_min, y_max - y_min))
    out = {}
    for i in range(n):
        warped = cv2.warpAffine(strips[i], Ms[i], sizes[i], flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
        content = CR.content_mask_cv2(strips[i], Ms[i], sizes[i])
        if gains[i] != (1.0, 1.0, 1.0):   # applyChannelGainInPlace: float32 multiply, convertTo(CV_8U)
            g = np.asarray(gains[i], np.float32)
            warped = np.clip(np.rint(warped.astype(np.float32) * g[None, None, :]), 0, 255).astype(np.uint8)
        seam = cv2.resize(seams[i], sizes[i], interpolation=cv2.INTER_NEAREST)
        _, seam = cv2.threshold(seam, 1.0, 255.0, cv2.THRESH_BINARY)
        soft = CR.soft_blend_mask_cv2(seam, content, 10.0)
        blender.feed(warped.astype(np.int16), soft, corners[i])
        out[f"strip{i}"] = strips[i]; out[f"M{i}"] = Ms[i]; out[f"seam{i}"] = seams[i]
        out[f"content{i}"] = content; out[f"soft{i}"] = soft; out[f"warped{i}"] = warped
    res, res_mask = blender.blend(None, None)
    pano = np.clip(res, 0, 255).astype(np.uint8)
    rect, _ = CR.auto_crop_rect_cv2(pano)
    np.savez_compressed(os.path.join(HERE, "global_stage.npz"), n=n, bands=bands, corners=np.array(corners), sizes=np.array(sizes),
                        gains=np.array(gains, np.float32), pano=pano, mask=res_mask, crop=np.array(rect),
                        roi=np.array([x_min, y_min, x_max - x_min, y_max - y_min]), cv_version=cv2.__version__, **out)
    print("global_stage", pano.shape, rect)


if __name__ == "__main__":
    global_stage_case()
    plane_case("compose_mb5", 3, 2, 200, 150, 0.6, 31, 0.4, "multiband", 5)
    plane_case("compose_mb3", 2, 2, 180, 140, 0.5, 32, 1.0, "multiband", 3)
    plane_case("compose_feather", 2, 2, 200, 150, 0.6, 33, 0.37, "feather", 0)
    pyramid_case()
    affine_case()
