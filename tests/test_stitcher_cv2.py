"""The drop-in claim at the reference's first call site, against the real thing: cv::Stitcher (SCANS mode, as
/root/reference/src/stitch_robust.cpp:178 creates it) registers a small synthetic strip and composes it; the library,
fed with what the stitcher holds after estimateTransform (:251) - cameras(), workScale() - plus the seam masks of the
stitcher's own seam finder, must produce the panorama composePanorama (:256) produces, bit for bit.

Everything the stitcher does between registration and blending is followed through OpenCV itself (cv2 calls in the order
of Stitcher::composePanorama): seam-scale warps, GraphCutSeamFinder, full-resolution warps, dilate + INTER_LINEAR_EXACT
resize of the seam masks, MultiBandBlender(5). That restatement is first checked against the real composePanorama
(SURVEY probe P19), then the library against both. Components the Python Stitcher cannot be given (the reference's
DpSeamFinder / BlocksGain, stitch_robust.cpp:207-211) are covered by parity_cases.case_seam_* / case_exposure_gains.
"""
import math

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from drone_image_stitch_cpp_b200 import compositor as CP  # noqa: E402
from helpers import assert_blend_parity  # noqa: E402


def rich_ortho(h, w, seed):
    """Corner-rich synthetic ground (the feature finder needs more than smooth noise)."""
    rng = np.random.default_rng(seed)
    img = rng.integers(60, 170, (h // 8 + 2, w // 8 + 2, 3)).astype(np.uint8)
    img = cv2.resize(img, (w, h), interpolation=cv2.INTER_CUBIC)
    for _ in range((h * w) // 1500):
        x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
        c = tuple(int(v) for v in rng.integers(0, 256, 3))
        k = rng.integers(0, 3)
        if k == 0:
            cv2.rectangle(img, (x, y), (x + int(rng.integers(4, 30)), y + int(rng.integers(4, 30))), c, -1)
        elif k == 1:
            cv2.circle(img, (x, y), int(rng.integers(3, 14)), c, -1)
        else:
            cv2.line(img, (x, y), (x + int(rng.integers(-40, 40)), y + int(rng.integers(-40, 40))), c, int(rng.integers(1, 4)))
    noise = rng.integers(-6, 7, img.shape)
    return np.clip(img.astype(int) + noise, 0, 255).astype(np.uint8)


def strip(seed=7, n=3, fw=640, fh=480):
    ortho = rich_ortho(fh + 220, 80 + fw + (n - 1) * 260 + 80, seed)
    rng = np.random.default_rng(seed + 1)
    imgs = []
    for i in range(n):
        # small rotation + translation per frame, cut with cv2 so overlaps are consistent
        th = math.radians(float(rng.uniform(-1.5, 1.5)))
        x0, y0 = 40 + i * 260 + float(rng.uniform(-8, 8)), 90 + float(rng.uniform(-15, 15))
        M = np.array([[math.cos(th), -math.sin(th), x0], [math.sin(th), math.cos(th), y0]], np.float64)
        imgs.append(cv2.warpAffine(ortho, M, (fw, fh), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP))
    return imgs


def real_stitcher(imgs):
    st = cv2.Stitcher.create(cv2.Stitcher_SCANS)
    status = st.estimateTransform(imgs)
    assert status == 0, f"estimateTransform failed: {status}"
    status, pano = st.composePanorama()
    assert status == 0
    return st, pano


def seam_phase_cv2(st, imgs):
    """The seam phase of Stitcher::composePanorama through cv2: -> (seam masks at seam scale, their corners)."""
    cams, ws = st.cameras(), st.workScale()
    h, w = imgs[0].shape[:2]
    seam_scale = min(1.0, math.sqrt(st.seamEstimationResol() * 1e6 / (w * h)))
    swa = seam_scale / ws
    focals = sorted(c.focal for c in cams)
    wis = focals[len(focals) // 2] if len(focals) % 2 else (focals[len(focals) // 2 - 1] + focals[len(focals) // 2]) * 0.5
    warper = cv2.PyRotationWarper("affine", float(np.float32(wis * swa)))
    corners, warped_f, masks = [], [], []
    for img, c in zip(imgs, cams):
        small = cv2.resize(img, None, fx=seam_scale, fy=seam_scale, interpolation=cv2.INTER_LINEAR_EXACT)
        K = c.K().astype(np.float32)
        K[0, 0] *= np.float32(swa); K[0, 2] *= np.float32(swa); K[1, 1] *= np.float32(swa); K[1, 2] *= np.float32(swa)
        R = np.ascontiguousarray(c.R, np.float32)
        corner, wimg = warper.warp(small, K, R, cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        _, wmask = warper.warp(np.full(small.shape[:2], 255, np.uint8), K, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        corners.append(corner)
        warped_f.append(wimg.astype(np.float32))
        masks.append(wmask)
    finder = cv2.detail_GraphCutSeamFinder("COST_COLOR")
    masks = finder.find(warped_f, corners, masks)
    return [np.ascontiguousarray(m.get() if hasattr(m, "get") else m) for m in masks], corners, (seam_scale, swa, wis)


def compose_phase_cv2(st, imgs, seam_masks):
    """The compose phase of Stitcher::composePanorama through cv2 (compose_scale = 1)."""
    cams, ws = st.cameras(), st.workScale()
    cwa = 1.0 / ws
    focals = sorted(c.focal for c in cams)
    wis = focals[len(focals) // 2] if len(focals) % 2 else (focals[len(focals) // 2 - 1] + focals[len(focals) // 2]) * 0.5
    warper = cv2.PyRotationWarper("affine", float(np.float32(wis * cwa)))
    Ks, Rs, corners, sizes = [], [], [], []
    for img, c in zip(imgs, cams):
        focal, ppx, ppy = c.focal * cwa, c.ppx * cwa, c.ppy * cwa
        K = np.array([[focal, 0, ppx], [0, focal * c.aspect, ppy], [0, 0, 1]], np.float64).astype(np.float32)
        R = np.ascontiguousarray(c.R, np.float32)
        roi = warper.warpRoi((img.shape[1], img.shape[0]), K, R)
        Ks.append(K); Rs.append(R); corners.append((roi[0], roi[1])); sizes.append((roi[2], roi[3]))
    bl = cv2.detail_MultiBandBlender(0, 5)
    bl.prepare(cv2.detail.resultRoi(corners, sizes))
    for img, K, R, corner, sm in zip(imgs, Ks, Rs, corners, seam_masks):
        _, wimg = warper.warp(img, K, R, cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        _, wmask = warper.warp(np.full(img.shape[:2], 255, np.uint8), K, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        dil = cv2.dilate(sm, None)
        up = cv2.resize(dil, (wmask.shape[1], wmask.shape[0]), interpolation=cv2.INTER_LINEAR_EXACT)
        bl.feed(wimg.astype(np.int16), cv2.bitwise_and(up, wmask), corner)
    res, res_mask = bl.blend(None, None)
    return np.clip(res, 0, 255).astype(np.uint8), res_mask


def _stitcher_case(lib):
    imgs = strip()
    st, pano_real = real_stitcher(imgs)
    seam_masks, _, (seam_scale, swa, wis) = seam_phase_cv2(st, imgs)
    pano_cv, mask_cv = compose_phase_cv2(st, imgs, seam_masks)
    # P19: the restated loop is composePanorama
    assert pano_cv.shape == pano_real.shape and np.array_equal(pano_cv, pano_real), "the cv2 restatement is not composePanorama"
    # the library, fed with the stitcher's state
    pano, mask, roi = CP.compose_panorama_from_stitcher(imgs, st.cameras(), st.workScale(), bands=5, seam_lowres=seam_masks, lib=lib)
    assert pano.shape == pano_real.shape
    assert np.array_equal(mask, mask_cv)
    assert_blend_parity(pano, pano_real)
    assert np.array_equal(pano, pano_real)
    # the seam-phase warps themselves (ds_warp_frame at seam scale) are the stitcher's
    c0 = st.cameras()[0]
    small = cv2.resize(imgs[0], None, fx=seam_scale, fy=seam_scale, interpolation=cv2.INTER_LINEAR_EXACT)
    K = c0.K().astype(np.float32)
    K[0, 0] *= np.float32(swa); K[0, 2] *= np.float32(swa); K[1, 1] *= np.float32(swa); K[1, 2] *= np.float32(swa)
    warper = cv2.PyRotationWarper("affine", float(np.float32(wis * swa)))
    corner, wimg = warper.warp(small, K, np.ascontiguousarray(c0.R, np.float32), cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
    got_corner, got_img, got_mask = CP.warp_frame(small, CP.plane_transform(K, c0.R, np.float32(wis * swa)), lib=lib)
    assert tuple(got_corner) == tuple(corner) and np.array_equal(got_img, wimg)


def test_real_stitcher_emu(emu_lib):
    _stitcher_case(emu_lib)


@pytest.mark.gpu
def test_real_stitcher_gpu(cuda_lib):
    _stitcher_case(cuda_lib)
