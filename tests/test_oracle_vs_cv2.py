"""Pins the C restatement (oracle/ds_oracle.c) against the real OpenCV (cv2 wheel) — the library the
reference's hot path runs in (/root/reference/CMakeLists.txt:18). CPU only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import cv_reference as CR  # noqa: E402
from oracle import ds_oracle as O  # noqa: E402

SIZES = [(2, 2), (4, 6), (32, 32), (64, 96), (33, 47), (128, 160), (5, 9), (96, 224), (20, 36), (8, 12), (40, 4), (4, 40),
         (16, 16), (6, 6), (10, 10), (12, 12), (14, 14), (18, 22), (1, 8), (8, 1), (3, 3)]


@pytest.mark.parametrize("hw", SIZES)
def test_pyramids(hw):
    h, w = hw
    rng = np.random.default_rng(h * 1000 + w)
    a = rng.integers(-300, 300, (h, w, 3)).astype(np.int16)
    assert np.array_equal(O.pyrdown_16s(a), cv2.pyrDown(a))
    assert np.array_equal(O.pyrup_16s(a), cv2.pyrUp(a))
    f = rng.random((h, w)).astype(np.float32)
    assert np.array_equal(O.pyrdown_f32(f), cv2.pyrDown(f))


def test_pyrdown_f32_chain_of_binary_mask():
    # weights of a binary mask pushed down 8 levels stay bit-exact (op order matters from level 4 on)
    m = np.zeros((768, 1024), np.float32)
    m[100:600, 200:900] = np.float32(255) * np.float32(1.0 / 255.0)
    a, b = m, m
    for _ in range(8):
        a, b = O.pyrdown_f32(a), cv2.pyrDown(b)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("ws,seed", [(1.0, 1), (0.37, 2), (0.45, 3)])
def test_maps_warp_mask(ws, seed):
    from drone_image_stitch_cpp_b200 import synth
    sv = synth.grid_survey(2, 2, 320, 240, overlap=0.6, seed=seed, work_scale=ws)
    for f, K, R in zip(sv.frames, sv.Ks, sv.Rs):
        roi, xm, ym, xy, a = CR.maps_cv2((f.shape[1], f.shape[0]), K, R, sv.scale)
        w = O.warp_frame(f, K, R, sv.scale)
        assert (roi[0], roi[1]) == w["corner"]
        assert np.array_equal(xy, w["xy"]) and np.array_equal(a, w["a"])   # INTER_BITS tables, bit-exact
        c, wi, mk = CR.warp_frame_cv2(f, K, R, sv.scale)
        assert tuple(c) == w["corner"]
        assert np.array_equal(wi, w["warped"])
        assert np.array_equal(mk, w["mask"])


def test_general_K_projector():
    # K with principal point and non-unit aspect: pins the float32 matmul order of the projector setup
    rng = np.random.default_rng(7)
    for _ in range(25):
        a = np.float32(rng.uniform(1.5, 4))
        K = np.array([[a, 0, rng.uniform(-50, 50)], [0, a * rng.uniform(0.9, 1.1), rng.uniform(-50, 50)], [0, 0, 1]], np.float32)
        th, s = rng.uniform(-0.1, 0.1), rng.uniform(0.9, 1.1)
        H = np.array([[s * np.cos(th), -s * np.sin(th), rng.uniform(-300, 300)],
                      [s * np.sin(th), s * np.cos(th), rng.uniform(-300, 300)], [0, 0, 1]], np.float32)
        for affine in (True, False):
            roi, xm, ym, xy, aa = CR.maps_cv2((200, 150), K, H, float(a), affine)
            k_rinv, r_kinv, t = O.projector_setup(K, H, affine)
            tl = O.plane_roi(r_kinv, t, float(a), 200, 150)
            assert (tl[0], tl[1], tl[2] - tl[0], tl[3] - tl[1]) == tuple(roi)
            x2, y2 = O.plane_maps(k_rinv, t, float(a), tl[0], tl[1], tl[2] - tl[0] + 1, tl[3] - tl[1] + 1)
            assert np.array_equal(x2, xm) and np.array_equal(y2, ym)


@pytest.mark.parametrize("blend,bands", [("multiband", 5), ("multiband", 3), ("multiband", 8), ("multiband", 1), ("feather", 0)])
def test_compose_matches_opencv(small_survey, blend, bands):
    sv = small_survey
    p1, m1, r1 = O.compose_port(sv.frames, sv.Ks, sv.Rs, sv.scale, blend, bands)
    p2, m2, r2 = CR.compose_cv2(sv.frames, sv.Ks, sv.Rs, sv.scale, blend, bands)
    assert r1 == r2
    assert np.array_equal(p1, p2)
    assert np.array_equal(m1, m2)


def test_warp_affine_and_perspective_tables():
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, (300, 400, 3)).astype(np.uint8)
    for _ in range(6):
        th, s = rng.uniform(-0.3, 0.3), rng.uniform(0.8, 1.2)
        M = np.array([[s * np.cos(th), -s * np.sin(th), rng.uniform(-50, 50)], [s * np.sin(th), s * np.cos(th), rng.uniform(-50, 50)]])
        dw, dh = 450, 380
        xy, a = O.affine_tables(M, dw, dh)
        assert np.array_equal(O.remap_bilinear(src, xy, a, "constant"),
                              cv2.warpAffine(src, M, (dw, dh), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT))
        H = np.vstack([M, [rng.uniform(-1e-4, 1e-4), rng.uniform(-1e-4, 1e-4), 1]])
        xy, a = O.persp_tables(H, dw, dh)
        assert np.array_equal(O.remap_bilinear(src, xy, a, "constant"),
                              cv2.warpPerspective(src, H, (dw, dh), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT))


def test_feather_weight_map():
    rng = np.random.default_rng(5)
    m = np.full((200, 300), 255, np.uint8)
    m[:30, :] = 0
    m[:, 250:] = 0
    m[100:120, 100:130] = 0
    ref = cv2.detail.createWeightMap(m, 0.02, None) if hasattr(cv2.detail, "createWeightMap") else None
    if ref is None:
        d = cv2.distanceTransform(m, cv2.DIST_L1, 3)
        ref = np.minimum(d * np.float32(0.02), np.float32(1.0)).astype(np.float32)
    assert np.array_equal(O.feather_weight_map(m, 0.02), np.asarray(ref))
    full = np.full((40, 50), 255, np.uint8)
    assert np.all(O.feather_weight_map(full, 0.02) == 1.0)


def test_degenerate_perspective_maps():
    # z crosses zero inside the frame: x/z overflows; cvRound's x86 "integer indefinite" must be reproduced
    K = np.array([[1.0, 0, 3.5], [0, 1.02, -2.25], [0, 0, 1]], np.float32)
    R = np.array([[0.9998, -0.019, 0.01], [0.019, 0.9998, -0.02], [1e-5, -2e-5, 1.0]], np.float32)
    img = np.random.default_rng(0).integers(0, 256, (180, 240, 3)).astype(np.uint8)
    roi, xm, ym, xy, a = CR.maps_cv2((240, 180), K, R, 1.0, affine=False)
    w = O.warp_frame(img, K, R, 1.0, affine=False)
    assert np.array_equal(xy, w["xy"]) and np.array_equal(a, w["a"])
    c, wi, mk = CR.warp_frame_cv2(img, K, R, 1.0, affine=False)
    assert np.array_equal(wi, w["warped"]) and np.array_equal(mk, w["mask"])


@pytest.mark.parametrize("shape", [(30, 40, 300, 410), (17, 23, 200, 333), (50, 60, 211, 257), (8, 9, 100, 90), (40, 30, 41, 31),
                                   (25, 31, 512, 640), (1, 7, 20, 70), (12, 12, 12, 12)])
def test_seam_mask_upsize(shape):
    # composePanorama: dilate(masks_warped[i]) -> resize(INTER_LINEAR_EXACT) -> AND   
    sh, sw, dh, dw = shape
    rng = np.random.default_rng(sh * 7 + sw)
    for src in ((rng.random((sh, sw)) > 0.5).astype(np.uint8) * 255, rng.integers(0, 256, (sh, sw)).astype(np.uint8)):
        ref = cv2.resize(cv2.dilate(src, None), (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT)
        assert np.array_equal(O.seam_mask_upsize(src, dw, dh), ref)


def test_gain_apply_semantics():
    """The three gain multiplies between warp and blend: float32 channel gain (stitch_global.cpp:291-305),
    float64 scalar gains of ExposureCompensator::apply (cv::multiply by a Scalar), float32 per-pixel map of
    BlocksGainCompensator::apply (cv::multiply 8UC3 x 32FC3 -> 8U). parity_cases.oracle_warp uses these forms."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (120, 160, 3)).astype(np.uint8)
    for g in [(1.07, 0.93, 1.21), (1.003, 0.997, 1.1)]:
        ref = cv2.multiply(img, np.array(g + (0,), np.float64))
        assert np.array_equal(ref, np.clip(np.rint(img.astype(np.float64) * np.array(g)), 0, 255).astype(np.uint8))
        f32 = img.astype(np.float32)
        chans = cv2.split(f32)
        chans = [c * np.float32(gg) for c, gg in zip(chans, g)]
        ref32 = cv2.convertScaleAbs(cv2.merge(chans))   # == convertTo(CV_8U) for non-negative data: cvRound + saturate
        assert np.array_equal(ref32, np.clip(np.rint(f32 * np.array(g, np.float32)), 0, 255).astype(np.uint8))
    gm = (rng.random((120, 160)) * 0.4 + 0.8).astype(np.float32)
    ref = cv2.multiply(img, cv2.merge([gm, gm, gm]), dtype=cv2.CV_8U)
    assert np.array_equal(ref, np.clip(np.rint(img.astype(np.float32) * gm[:, :, None]), 0, 255).astype(np.uint8))


# ---- the global stage's masks (src/stitch_global.cpp:328-383, :649-655)

def _strip_like(rng, h, w, holes=True):
    """A strip panorama stand-in: textured content with black wedges / holes (what autoCrop leaves behind)."""
    img = rng.integers(8, 256, (h, w, 3), dtype=np.uint8)
    if holes:
        yy, xx = np.mgrid[0:h, 0:w]
        img[(yy < 0.15 * xx - 10) | (yy > h - 12 + 0.05 * xx)] = 0
        img[h // 3:h // 3 + 17, w // 2:w // 2 + 40] = rng.integers(0, 5, (17, 40, 3), dtype=np.uint8)   # near-black noise
    return img


def test_bgr2gray():
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (211, 317, 3), dtype=np.uint8)
    img[:40] = rng.integers(0, 9, (40, 317, 3), dtype=np.uint8)
    assert np.array_equal(O.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("seed", range(6))
def test_content_mask(seed):
    rng = np.random.default_rng(100 + seed)
    sh, sw = 180 + 17 * seed, 260 + 11 * seed
    img = _strip_like(rng, sh, sw, holes=seed % 3 != 2)
    ang, s = np.deg2rad(rng.uniform(-25, 25)), rng.uniform(0.8, 1.3)
    M = np.array([[s * np.cos(ang), -s * np.sin(ang), rng.uniform(-20, 60)], [s * np.sin(ang), s * np.cos(ang), rng.uniform(-20, 60)]])
    if seed == 4:
        M = np.array([[1.0, 0, 5.03125], [0, 1.0, 3.03125]])   # fractions of 1/32: single taps of weight 1/1024 decide
    dw, dh = 390, 310
    assert np.array_equal(O.content_mask(img, M, dw, dh), CR.content_mask_cv2(img, M, (dw, dh)))


@pytest.mark.parametrize("dims", [(37, 29, 400, 330), (123, 77, 1001, 613), (50, 50, 50, 50), (64, 48, 63, 47), (7, 5, 1000, 999)])
def test_resize_nearest(dims):
    sw, sh, dw, dh = dims
    m = np.random.default_rng(sw).integers(0, 2, (sh, sw), dtype=np.uint8) * 255
    assert np.array_equal(O.resize_nearest(m, dw, dh), cv2.resize(m, (dw, dh), interpolation=cv2.INTER_NEAREST))


def test_gaussian_kernel():
    for sigma in (10.0, 3.0, 7.5, 1.0):
        n = int(np.rint(sigma * 8 + 1)) | 1
        assert np.array_equal(O.gaussian_kernel_f32(n, sigma), cv2.getGaussianKernel(n, sigma, cv2.CV_32F).ravel())


@pytest.mark.parametrize("hw,sigma", [((300, 421), 10.0), ((257, 256), 10.0), ((90, 77), 10.0), ((33, 500), 10.0), ((640, 37), 10.0),
                                      ((301, 333), 4.0), ((1200, 1611), 10.0)])
def test_soft_blend_mask(hw, sigma):
    h, w = hw
    rng = np.random.default_rng(h + w)
    seam = (cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (0, 0), 25) > 0.5).astype(np.uint8) * 255
    content = np.full((h, w), 255, np.uint8)
    content[:, :5] = 0
    content[h // 2:h // 2 + 9, w // 3:w // 3 + 30] = 0
    got = O.soft_blend_mask(seam, content, sigma)
    ref = CR.soft_blend_mask_cv2(seam, content, sigma)
    assert np.array_equal(got, ref)
    assert ((got > 0) & (got < 255)).any() and (got == 0).any()


# ---- autoCropBlackBorder (src/stitch_common.cpp:4-27)

def _pano_like(rng, h, w, blobs, speck=0):
    img = np.zeros((h, w, 3), np.uint8)
    m = np.zeros((h, w), np.float32)
    for _ in range(blobs):
        cy, cx = rng.integers(0, h), rng.integers(0, w)
        ry, rx = rng.integers(3, max(4, h // 3)), rng.integers(3, max(4, w // 3))
        yy, xx = np.mgrid[0:h, 0:w]
        m[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1] = 1
    img[m > 0] = rng.integers(2, 256, (int((m > 0).sum()), 3), dtype=np.uint8)
    for _ in range(speck):
        img[rng.integers(0, h), rng.integers(0, w)] = rng.integers(0, 256, 3)
    hole = rng.random((h, w)) < 0.02
    img[hole] = rng.integers(0, 2, (int(hole.sum()), 3), dtype=np.uint8)   # dark pixels inside the content
    return img


@pytest.mark.parametrize("seed", range(12))
def test_auto_crop_rect(seed):
    rng = np.random.default_rng(900 + seed)
    h, w = int(rng.integers(20, 90)), int(rng.integers(20, 120))
    img = _pano_like(rng, h, w, blobs=int(rng.integers(1, 5)), speck=int(rng.integers(0, 6)))
    rect, areas = CR.auto_crop_rect_cv2(img)
    srt = sorted(areas)
    if len(srt) > 1 and srt[-1] == srt[-2]:
        pytest.skip("tie between two maximal contours: order of cv::findContours decides")
    assert O.auto_crop_rect(img) == rect


def test_auto_crop_rect_empty_and_full():
    z = np.zeros((17, 23, 3), np.uint8)
    assert O.auto_crop_rect(z) == CR.auto_crop_rect_cv2(z)[0] == (0, 0, 23, 17)
    z[:] = 200
    assert O.auto_crop_rect(z) == CR.auto_crop_rect_cv2(z)[0] == (0, 0, 23, 17)


@pytest.mark.parametrize("sw,sh,dw,dh", [(23, 17, 731, 540), (171, 114, 1368, 912), (9, 7, 280, 210), (40, 30, 40, 30), (5, 3, 161, 97),
                                          (2, 2, 64, 64), (2, 3, 20, 20), (60, 45, 1900, 1430)])
def test_resize_linear_f32(sw, sh, dw, dh):
    """cv::resize(CV_32FC1, INTER_LINEAR) - the gain-map upsizing of BlocksGainCompensator::apply
    (/root/reference/src/stitch_robust.cpp:209-211) - pinned bit for bit (round-1 left it with the caller). Maps with a
    single row or column (frames under 33 px in one dimension) take another path inside OpenCV and are not pinned: the
    library refuses them (DS_ERR_UNSUPPORTED) and the caller passes the resized map as gain_map."""
    rng = np.random.default_rng(sw * 1000 + dh)
    src = (1.0 + 0.3 * rng.standard_normal((sh, sw))).astype(np.float32)
    ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    got = O.resize_linear_f32(src, dw, dh)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_blocks_gain_apply_chain():
    """BlocksGainCompensator.apply == sat_u8(rint(img * resize_linear_f32(gain blocks))), through the
    real compensator: feed it two overlapping images, read its gain maps back, apply, and restate."""
    rng = np.random.default_rng(5)
    h, w = 200, 300
    base = rng.integers(30, 200, (h, w, 3)).astype(np.uint8)
    imgs = [base.copy(), np.clip(base.astype(np.float32) * 1.25, 0, 255).astype(np.uint8)]
    masks = [np.full((h, w), 255, np.uint8), np.full((h, w), 255, np.uint8)]
    comp = cv2.detail_BlocksGainCompensator(32, 32)
    comp.feed([(0, 0), (0, 0)], imgs, masks)
    maps = comp.getMatGains()
    for i in range(2):
        gm = np.asarray(maps[i].get() if hasattr(maps[i], "get") else maps[i], np.float32)
        assert gm.shape == ((h + 31) // 32, (w + 31) // 32)
        want = imgs[i].copy()
        comp.apply(i, (0, 0), want, masks[i])
        full = O.resize_linear_f32(gm, w, h)
        got = np.clip(np.rint(imgs[i].astype(np.float32) * full[:, :, None]), 0, 255).astype(np.uint8)
        assert np.array_equal(got, want)
