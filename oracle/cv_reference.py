"""The reference's CPU path, run through the real OpenCV (cv2 wheel).

TEST INFRASTRUCTURE ONLY (see oracle/ds_oracle.c header). This module does not restate anything:
it calls the very OpenCV classes the reference configures
(/root/reference/src/stitch_robust.cpp:203-213: AffineWarper, MultiBandBlender(bands); image warp
INTER_LINEAR + BORDER_REFLECT and mask warp INTER_NEAREST + BORDER_CONSTANT are what
cv::Stitcher::composePanorama, called at :256, does per frame) in the order composePanorama calls
them. It is the "reference" arm of bench.py and the generator of tests/golden/.
The C++ reference itself cannot be built here (no OpenCV C++ SDK, CMakeLists.txt:18).
"""
import os

import numpy as np


def have_cv2():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def set_threads(n=None):
    import cv2
    cv2.setNumThreads(int(n if n else (os.cpu_count() or 1)))
    return cv2.getNumThreads()


def warp_frame_cv2(img, K, R, scale, affine=True):
    import cv2
    w = cv2.PyRotationWarper("affine" if affine else "plane", float(scale))
    K = np.ascontiguousarray(K, np.float32)
    R = np.ascontiguousarray(R, np.float32)
    corner, warped = w.warp(img, K, R, cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
    mask = np.full(img.shape[:2], 255, np.uint8)
    _, mask_w = w.warp(mask, K, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
    return corner, warped, mask_w


def maps_cv2(size_wh, K, R, scale, affine=True):
    """-> roi(x,y,w,h as returned by buildMaps), xy int16x2, a uint16 (cv2.convertMaps)."""
    import cv2
    w = cv2.PyRotationWarper("affine" if affine else "plane", float(scale))
    roi, xm, ym = w.buildMaps(size_wh, np.ascontiguousarray(K, np.float32), np.ascontiguousarray(R, np.float32))
    xy, a = cv2.convertMaps(xm, ym, cv2.CV_16SC2)
    return roi, xm, ym, xy, a


def compose_cv2(frames, Ks, Rs, scale, blend="multiband", bands=5, sharpness=0.02, affine=True, timings=None):
    """Same contract as ds_oracle.compose_port, executed by OpenCV itself."""
    import time

    import cv2
    t0 = time.perf_counter()
    warper = cv2.PyRotationWarper("affine" if affine else "plane", float(scale))
    corners, sizes = [], []
    for f, K, R in zip(frames, Ks, Rs):
        roi = warper.warpRoi((f.shape[1], f.shape[0]), np.ascontiguousarray(K, np.float32),
                             np.ascontiguousarray(R, np.float32))
        corners.append((roi[0], roi[1]))
        sizes.append((roi[2], roi[3]))
    if blend == "multiband":
        bl = cv2.detail_MultiBandBlender(0, int(bands))
    else:
        bl = cv2.detail_FeatherBlender(float(sharpness))
    roi = cv2.detail.resultRoi(corners, sizes)
    bl.prepare(roi)
    for f, K, R, c in zip(frames, Ks, Rs, corners):
        K = np.ascontiguousarray(K, np.float32)
        R = np.ascontiguousarray(R, np.float32)
        _, warped = warper.warp(f, K, R, cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        mask = np.full(f.shape[:2], 255, np.uint8)
        _, mask_w = warper.warp(mask, K, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        bl.feed(warped.astype(np.int16), mask_w, c)
    res, res_mask = bl.blend(None, None)
    pano = cv2.convertScaleAbs(res) if False else np.clip(res, 0, 255).astype(np.uint8)
    if timings is not None:
        timings["seconds"] = time.perf_counter() - t0
    return pano, res_mask, tuple(int(v) for v in roi)


# ---- the global stage's mask helpers, through cv2 itself (src/stitch_global.cpp:328-383)

def content_mask_cv2(img, M, dsize):
    """buildWarpedContentMask."""
    import cv2
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    _, m8 = cv2.threshold(gray, 3, 255, cv2.THRESH_BINARY)
    mf = cv2.multiply(m8, 1.0 / 255.0, dtype=cv2.CV_32F)   # convertTo(CV_32F, 1/255): 255 -> exactly 1.0f either way
    wf = cv2.warpAffine(mf, np.asarray(M, np.float64).reshape(2, 3), dsize, flags=cv2.INTER_LINEAR,
                        borderMode=cv2.BORDER_CONSTANT, borderValue=0.0)
    _, mw = cv2.threshold(wf, 0.999, 255.0, cv2.THRESH_BINARY)
    return mw.astype(np.uint8)


def soft_blend_mask_cv2(seam, content, sigma=10.0):
    """buildSoftBlendMask."""
    import cv2
    b = cv2.bitwise_and(seam, content)
    _, b = cv2.threshold(b, 1.0, 255.0, cv2.THRESH_BINARY)
    bf = (b > 0).astype(np.float32)                        # convertTo(CV_32F, 1/255) of a 0 / 255 image
    soft = cv2.GaussianBlur(bf, (0, 0), sigma, None, sigma, cv2.BORDER_REPLICATE)
    soft = cv2.multiply(soft, bf)
    # convertTo(CV_8U, 255.0) scales in float32 (cvt_32f); cv2.normalize(MINMAX 0..255) issues exactly that call when the
    # input spans [0, 1] - here done directly
    return np.clip(np.rint((soft * np.float32(255.0)).astype(np.float32)), 0, 255).astype(np.uint8)


def auto_crop_rect_cv2(pano):
    """autoCropBlackBorder (src/stitch_common.cpp:4-27) through cv2; -> (x, y, w, h), plus the contour areas."""
    import cv2
    gray = cv2.cvtColor(pano, cv2.COLOR_BGR2GRAY)
    _, th = cv2.threshold(gray, 1, 255, cv2.THRESH_BINARY)
    contours, _ = cv2.findContours(th, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    if not contours:
        return (0, 0, pano.shape[1], pano.shape[0]), []
    areas = [cv2.contourArea(c) for c in contours]
    best, best_area = cv2.boundingRect(contours[0]), areas[0]
    for c, a in zip(contours[1:], areas[1:]):
        if a > best_area:
            best_area, best = a, cv2.boundingRect(c)
    return tuple(int(v) for v in best), areas
