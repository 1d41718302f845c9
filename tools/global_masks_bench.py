"""Global-stage mask chain at strip-panorama size: device path through the C ABI vs the
reference's own OpenCV calls (buildWarpedContentMask + NEAREST seam resize + buildSoftBlendMask) on the host cores.
Usage: python tools/global_masks_bench.py [strip_w strip_h n_strips]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from drone_image_stitch_cpp_b200 import compositor as CP  # noqa: E402

sw, sh, n = (int(v) for v in (sys.argv[1:4] + ["12000", "4000", "3"][len(sys.argv) - 1:]))
rng = np.random.default_rng(5)
strips, Ms, rois, seams = [], [], [], []
for i in range(n):
    img = rng.integers(8, 256, (sh, sw, 3), dtype=np.uint8)
    yy = np.arange(sh)[:, None]
    xx = np.arange(sw)[None, :]
    img[(yy < 0.02 * xx - 40) | (yy > sh - 60 + 0.01 * xx)] = 0
    th = np.deg2rad(1.0 + i)
    Hm = np.array([[np.cos(th), -np.sin(th), 300.0], [np.sin(th), np.cos(th), 50.0 + i * sh * 0.6]])
    pts = Hm @ np.array([[0, sw, sw, 0], [0, 0, sh, sh], [1, 1, 1, 1]], np.float64)
    x0, y0 = int(np.floor(pts[0].min())), int(np.floor(pts[1].min()))
    bw, bh = int(np.ceil(pts[0].max())) - x0, int(np.ceil(pts[1].max())) - y0
    M = Hm.copy(); M[0, 2] -= x0; M[1, 2] -= y0
    strips.append(img); Ms.append(M); rois.append((x0, y0, bw, bh))
    low = np.full((bh // 8, bw // 8), 255, np.uint8)
    if i:
        low[: int(0.18 * low.shape[0])] = 0      # seam against the strip above (strips advance by 0.6 h: the kept parts meet)
    if i < n - 1:
        low[int(0.82 * low.shape[0]):] = 0       # and below
    seams.append(low)

roi = CP.result_roi(rois)
cv = CP.Canvas(roi, "multiband", 5)
xfs = [CP.affine_transform(M, r[:2], r[2:]) for M, r in zip(Ms, rois)]
t_up, t_masks = [], []
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(n):
        cv.upload(i, strips[i], xfs[i])
    t1 = time.perf_counter()
    for i in range(n):
        cv.update_opts(i, seam_lowres=seams[i], seam_nearest=True, content_mask=True, soft_mask=True)
    t2 = time.perf_counter()
    t_up.append(t1 - t0); t_masks.append(t2 - t1)
got = [cv.frame_mask(i, 0) for i in range(n)]
cv.set_profiling(True)
cv.composite()
kt = cv.kernel_times()
ms_comp = sum(k["ms"] for k in kt)

import cv2
cv2.setNumThreads(os.cpu_count())


# The reference's own OpenCV calls (src/stitch_global.cpp:353-383, :332-351; src/stitch_common.cpp:4-27), timed on the host
def content_mask_cv2(img, M, dsize):
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    _, m8 = cv2.threshold(gray, 3, 255, cv2.THRESH_BINARY)
    mf = cv2.multiply(m8, 1.0 / 255.0, dtype=cv2.CV_32F)
    wf = cv2.warpAffine(mf, np.asarray(M, np.float64).reshape(2, 3), dsize, flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0.0)
    _, mw = cv2.threshold(wf, 0.999, 255.0, cv2.THRESH_BINARY)
    return mw.astype(np.uint8)


def soft_blend_mask_cv2(seam, content, sigma=10.0):
    b = cv2.bitwise_and(seam, content)
    _, b = cv2.threshold(b, 1.0, 255.0, cv2.THRESH_BINARY)
    bf = cv2.multiply(b, 1.0 / 255.0, dtype=cv2.CV_32F)
    soft = cv2.GaussianBlur(bf, (0, 0), sigma, None, sigma, cv2.BORDER_REPLICATE)
    soft = cv2.multiply(soft, bf)
    return np.clip(np.rint(soft * np.float32(255.0)), 0, 255).astype(np.uint8)   # convertTo(CV_8U, 255.0): float32 scaling


def auto_crop_rect_cv2(pano):
    gray = cv2.cvtColor(pano, cv2.COLOR_BGR2GRAY)
    _, th = cv2.threshold(gray, 1, 255, cv2.THRESH_BINARY)
    contours, _ = cv2.findContours(th, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    if not contours:
        return (0, 0, pano.shape[1], pano.shape[0])
    best, best_area = cv2.boundingRect(contours[0]), cv2.contourArea(contours[0])
    for c in contours[1:]:
        a = cv2.contourArea(c)
        if a > best_area:
            best_area, best = a, cv2.boundingRect(c)
    return tuple(int(v) for v in best)


# autoCropBlackBorder: rectangle decided on the device vs cv2 on the downloaded panorama
t_crop = []
for rep in range(3):
    t0 = time.perf_counter()
    rect = cv.auto_crop_rect()
    t_crop.append(time.perf_counter() - t0)
t0 = time.perf_counter()
pano, _ = cv.download()
t_dl = time.perf_counter() - t0
t0 = time.perf_counter()
rect_cv = auto_crop_rect_cv2(pano)
t_crop_cpu = time.perf_counter() - t0
t0 = time.perf_counter()
ref = []
for i in range(n):
    bw, bh = rois[i][2], rois[i][3]
    content = content_mask_cv2(strips[i], Ms[i], (bw, bh))
    seam = cv2.resize(seams[i], (bw, bh), interpolation=cv2.INTER_NEAREST)
    _, seam = cv2.threshold(seam, 1.0, 255.0, cv2.THRESH_BINARY)
    ref.append(soft_blend_mask_cv2(seam, content, 10.0))
t_cpu = time.perf_counter() - t0
same = [bool(np.array_equal(a, b)) for a, b in zip(got, ref)]
mp = sum(r[2] * r[3] for r in rois) / 1e6
print(json.dumps({"strips": n, "strip": [sw, sh], "mask_megapixels": round(mp, 1), "device_masks_ms": round(min(t_masks) * 1e3, 2),
                  "device_masks_MP_per_s": round(mp / min(t_masks), 1), "upload_ms": round(min(t_up) * 1e3, 1),
                  "cv2_masks_ms": round(t_cpu * 1e3, 1), "cv2_threads": os.cpu_count(), "identical_to_cv2": same,
                  "composite_ms": round(ms_comp, 2), "canvas": [roi[2], roi[3]],
                  "crop_rect": list(rect), "crop_rect_cv2": list(rect_cv), "device_crop_ms": round(min(t_crop) * 1e3, 2),
                  "cv2_crop_ms": round(t_crop_cpu * 1e3, 1), "full_download_ms": round(t_dl * 1e3, 1)}))
