"""ctypes front-end of the CPU oracle (oracle/ds_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package never imports this module.

`compose_port` restates the per-frame loop of cv::Stitcher::composePanorama as the reference
configures it (/root/reference/src/stitch_robust.cpp:203-213, :256 — AffineWarper, INTER_LINEAR +
BORDER_REFLECT image warp, INTER_NEAREST + BORDER_CONSTANT mask warp, ->16S, MultiBandBlender feed /
blend, ->8U), with FeatherBlender(0.02) selectable for BASELINE config 1.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libds_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "ds_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_mbb_create.restype = C.c_void_p
        _lib.orc_feather_create.restype = C.c_void_p
        _lib.orc_mbb_num_bands.restype = C.c_int
        _lib.orc_get_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def set_threads(n):
    lib().orc_set_threads(C.c_int(int(n)))


def get_threads():
    return int(lib().orc_get_threads())


def projector_setup(K, R, affine=True):
    K = np.ascontiguousarray(K, np.float32).reshape(9)
    R = np.ascontiguousarray(R, np.float32).reshape(9)
    k_rinv = np.empty(9, np.float32)
    r_kinv = np.empty(9, np.float32)
    t = np.empty(3, np.float32)
    lib().orc_projector_setup(_p(K), _p(R), C.c_int(1 if affine else 0), _p(k_rinv), _p(r_kinv), _p(t))
    return k_rinv, r_kinv, t


def plane_roi(r_kinv, t, scale, w, h):
    """-> (tl_x, tl_y, br_x, br_y), br inclusive (PlaneWarper::detectResultRoi)."""
    roi = np.empty(4, np.int32)
    lib().orc_plane_roi(_p(r_kinv), _p(t), C.c_float(scale), C.c_int(w), C.c_int(h), _p(roi))
    return tuple(int(v) for v in roi)


def plane_maps(k_rinv, t, scale, tlx, tly, mw, mh):
    xm = np.empty((mh, mw), np.float32)
    ym = np.empty((mh, mw), np.float32)
    lib().orc_plane_maps(_p(k_rinv), _p(t), C.c_float(scale), C.c_int(tlx), C.c_int(tly), C.c_int(mw), C.c_int(mh),
                         _p(xm), _p(ym))
    return xm, ym


def fixed_tables(xm, ym):
    xm = np.ascontiguousarray(xm, np.float32)
    ym = np.ascontiguousarray(ym, np.float32)
    xy = np.empty(xm.shape + (2,), np.int16)
    a = np.empty(xm.shape, np.uint16)
    lib().orc_fixed_tables(_p(xm), _p(ym), C.c_size_t(xm.size), _p(xy), _p(a))
    return xy, a


def affine_tables(M, dw, dh):
    M = np.ascontiguousarray(M, np.float64).reshape(6)
    xy = np.empty((dh, dw, 2), np.int16)
    a = np.empty((dh, dw), np.uint16)
    lib().orc_affine_tables(_p(M), C.c_int(dw), C.c_int(dh), _p(xy), _p(a))
    return xy, a


def persp_tables(H, dw, dh):
    H = np.ascontiguousarray(H, np.float64).reshape(9)
    xy = np.empty((dh, dw, 2), np.int16)
    a = np.empty((dh, dw), np.uint16)
    lib().orc_persp_tables(_p(H), C.c_int(dw), C.c_int(dh), _p(xy), _p(a))
    return xy, a


def affine_nearest_mask(M, dw, dh, sw, sh):
    M = np.ascontiguousarray(M, np.float64).reshape(6)
    m = np.empty((dh, dw), np.uint8)
    lib().orc_affine_nearest_mask(_p(M), C.c_int(dw), C.c_int(dh), C.c_int(sw), C.c_int(sh), _p(m))
    return m


def persp_nearest_mask(H, dw, dh, sw, sh):
    H = np.ascontiguousarray(H, np.float64).reshape(9)
    m = np.empty((dh, dw), np.uint8)
    lib().orc_persp_nearest_mask(_p(H), C.c_int(dw), C.c_int(dh), C.c_int(sw), C.c_int(sh), _p(m))
    return m


def remap_bilinear(src, xy, a, border="reflect"):
    assert src.dtype == np.uint8 and src.ndim == 3 and src.shape[2] == 3 and src.strides[2] == 1 and src.strides[1] == 3
    dh, dw = a.shape
    dst = np.empty((dh, dw, 3), np.uint8)
    xy = np.ascontiguousarray(xy)
    a = np.ascontiguousarray(a)
    lib().orc_remap_bilinear_u8c3(_p(src), C.c_int(src.shape[1]), C.c_int(src.shape[0]), C.c_size_t(src.strides[0]),
                                  _p(xy), _p(a), C.c_int(dw), C.c_int(dh), _p(dst), C.c_size_t(dw * 3),
                                  C.c_int(1 if border == "reflect" else 0))
    return dst


def nearest_mask(xm, ym, sw, sh):
    xm = np.ascontiguousarray(xm, np.float32)
    ym = np.ascontiguousarray(ym, np.float32)
    m = np.empty(xm.shape, np.uint8)
    lib().orc_nearest_mask(_p(xm), _p(ym), C.c_size_t(xm.size), C.c_int(sw), C.c_int(sh), _p(m))
    return m


def pyrdown_16s(src):
    src = np.ascontiguousarray(src, np.int16)
    h, w = src.shape[:2]
    cn = 1 if src.ndim == 2 else src.shape[2]
    shape = ((h + 1) // 2, (w + 1) // 2) + (() if src.ndim == 2 else (cn,))
    dst = np.empty(shape, np.int16)
    lib().orc_pyrdown_16s(_p(src), C.c_int(w), C.c_int(h), C.c_int(cn), _p(dst))
    return dst


def pyrdown_f32(src):
    src = np.ascontiguousarray(src, np.float32)
    h, w = src.shape
    dst = np.empty(((h + 1) // 2, (w + 1) // 2), np.float32)
    lib().orc_pyrdown_f32(_p(src), C.c_int(w), C.c_int(h), _p(dst))
    return dst


def pyrup_16s(src):
    src = np.ascontiguousarray(src, np.int16)
    h, w = src.shape[:2]
    cn = 1 if src.ndim == 2 else src.shape[2]
    shape = (2 * h, 2 * w) + (() if src.ndim == 2 else (cn,))
    dst = np.empty(shape, np.int16)
    lib().orc_pyrup_16s(_p(src), C.c_int(w), C.c_int(h), C.c_int(cn), _p(dst))
    return dst


def feather_weight_map(mask, sharpness=0.02):
    mask = np.ascontiguousarray(mask, np.uint8)
    h, w = mask.shape
    out = np.empty((h, w), np.float32)
    lib().orc_feather_weight_map(_p(mask), C.c_int(w), C.c_int(h), C.c_float(sharpness), _p(out))
    return out


def seam_mask_upsize(mask, dw, dh):
    """dilate(3x3) + resize(INTER_LINEAR_EXACT) of a low-res seam mask to (dw, dh)."""
    mask = np.ascontiguousarray(mask, np.uint8)
    sh, sw = mask.shape
    out = np.empty((dh, dw), np.uint8)
    lib().orc_seam_mask_upsize(_p(mask), C.c_int(sw), C.c_int(sh), C.c_size_t(mask.strides[0]), C.c_int(dw), C.c_int(dh), _p(out))
    return out


def resize_linear_f32(src, dw, dh):
    """cv::resize(CV_32FC1, INTER_LINEAR) of the declared OpenCV build: BlocksGainCompensator::apply's gain-map upsizing."""
    src = np.ascontiguousarray(src, np.float32)
    sh, sw = src.shape
    out = np.empty((dh, dw), np.float32)
    lib().orc_resize_linear_f32(_p(src), C.c_int(sw), C.c_int(sh), C.c_size_t(src.strides[0] // 4), C.c_int(dw), C.c_int(dh), _p(out))
    return out


# ---- global-stage masks (src/stitch_global.cpp:328-383, :649-655)

def bgr2gray(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(img.shape[:2], np.uint8)
    lib().orc_bgr2gray(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_size_t(img.strides[0]), _p(out))
    return out


def content_mask(img, M, dw, dh):
    """buildWarpedContentMask(src_image, affine, dst_size)."""
    img = np.ascontiguousarray(img, np.uint8)
    M = np.ascontiguousarray(M, np.float64).reshape(-1)[:6].copy()
    out = np.empty((dh, dw), np.uint8)
    lib().orc_content_mask(_p(img), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_size_t(img.strides[0]), _p(M),
                           C.c_int(dw), C.c_int(dh), _p(out))
    return out


def resize_nearest(mask, dw, dh):
    mask = np.ascontiguousarray(mask, np.uint8)
    out = np.empty((dh, dw), np.uint8)
    lib().orc_resize_nearest_u8(_p(mask), C.c_int(mask.shape[1]), C.c_int(mask.shape[0]), C.c_size_t(mask.strides[0]),
                                C.c_int(dw), C.c_int(dh), _p(out))
    return out


def gaussian_kernel_f32(n, sigma):
    k = np.empty(n, np.float32)
    lib().orc_gaussian_kernel_f32(C.c_int(n), C.c_double(sigma), _p(k))
    return k


def soft_blend_mask(seam, content, sigma=10.0):
    """buildSoftBlendMask(seam_mask, content_mask); either may be None (= all 255)."""
    ref = seam if seam is not None else content
    h, w = ref.shape
    seam = None if seam is None else np.ascontiguousarray(seam, np.uint8)
    content = None if content is None else np.ascontiguousarray(content, np.uint8)
    out = np.empty((h, w), np.uint8)
    lib().orc_soft_blend_mask(_p(seam) if seam is not None else None, C.c_size_t(seam.strides[0] if seam is not None else 0),
                              _p(content) if content is not None else None,
                              C.c_size_t(content.strides[0] if content is not None else 0),
                              C.c_int(w), C.c_int(h), C.c_double(sigma), _p(out))
    return out


def threshold_gt1(mask):
    """ensureBinaryMask: threshold(mask, 1, 255, THRESH_BINARY)."""
    return np.where(np.asarray(mask) > 1, 255, 0).astype(np.uint8)


# ---- autoCropBlackBorder (src/stitch_common.cpp:4-27)

def _outer_contour(fg, i, j):
    """Suzuki-Abe border following of the outer border (8-connected foreground) that starts at the raster-first
    pixel (i, j) of a component; -> list of (x, y). Pure Python: small cases only."""
    h, w = fg.shape
    # 8-neighbourhood in clockwise order on the screen (y down): E, SE, S, SW, W, NW, N, NE
    nb = [(0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1)]

    def val(y, x):
        return 0 <= y < h and 0 <= x < w and fg[y, x]

    def idx(dy, dx):
        return nb.index((dy, dx))

    # step 3.1: clockwise from the west neighbour
    k0 = idx(0, -1)
    first = None
    for t in range(8):
        dy, dx = nb[(k0 + t) % 8]
        if val(i + dy, j + dx):
            first = (i + dy, j + dx)
            break
    if first is None:
        return [(j, i)]
    pts = []
    i2, j2 = first
    i3, j3 = i, j
    while True:
        # step 3.3: counter-clockwise around (i3, j3), starting after (i2, j2)
        k = idx(i2 - i3, j2 - j3)
        for t in range(1, 9):
            dy, dx = nb[(k - t) % 8]
            if val(i3 + dy, j3 + dx):
                i4, j4 = i3 + dy, j3 + dx
                break
        pts.append((j3, i3))
        if (i4, j4) == (i, j) and (i3, j3) == first:
            break
        i2, j2 = i3, j3
        i3, j3 = i4, j4
    return pts


def auto_crop_rect(pano):
    """The rectangle autoCropBlackBorder keeps: BGR2GRAY > 1, external contours (8-connected components not nested in
    another one's hole are what RETR_EXTERNAL returns; nested ones are smaller than their host and never win), the one of
    largest contourArea (shoelace over the border pixels; first wins a tie in raster order of the start pixels - OpenCV
    lists contours in reverse, a tie between distinct maxima is not pinned), its boundingRect. (x, y, w, h)."""
    from scipy import ndimage
    gray = bgr2gray(np.ascontiguousarray(pano[:, :, :3]))
    fg = gray > 1
    if not fg.any():
        return (0, 0, pano.shape[1], pano.shape[0])
    lbl, n = ndimage.label(fg, structure=np.ones((3, 3), int))
    best, best_area = None, -1.0
    for k, sl in enumerate(ndimage.find_objects(lbl), start=1):
        comp = lbl == k
        ys, xs = np.nonzero(comp[sl])
        i = int(ys.min())
        j = int(xs[ys == i].min())
        pts = _outer_contour(comp, i + sl[0].start, j + sl[1].start)
        x = np.array([p[0] for p in pts], np.float64)
        y = np.array([p[1] for p in pts], np.float64)
        area = abs(0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y)))
        if area > best_area:
            best_area = area
            best = (sl[1].start, sl[0].start, sl[1].stop - sl[1].start, sl[0].stop - sl[0].start)
    return best


def mbb_feed_geometry(roi, bands, tl, size):
    out = np.empty(8, np.int32)
    lib().orc_mbb_feed_geometry(C.c_int(roi[0]), C.c_int(roi[1]), C.c_int(roi[2]), C.c_int(roi[3]), C.c_int(bands),
                                C.c_int(tl[0]), C.c_int(tl[1]), C.c_int(size[0]), C.c_int(size[1]), _p(out))
    return [int(v) for v in out]


class MultiBand:
    """cv::detail::MultiBandBlender(false, bands, CV_32F) restated."""

    def __init__(self, roi, bands):
        x, y, w, h = roi
        self._h = C.c_void_p(lib().orc_mbb_create(C.c_int(x), C.c_int(y), C.c_int(w), C.c_int(h), C.c_int(bands)))
        info = np.empty(5, np.int32)
        lib().orc_mbb_info(self._h, _p(info))
        self.roi = tuple(int(v) for v in info[:4])  # padded
        self.bands = int(info[4])
        self.final = (w, h)

    def feed(self, img16, mask, tl, taps=False):
        img16 = np.ascontiguousarray(img16, np.int16)
        mask = np.ascontiguousarray(mask, np.uint8)
        ih, iw = mask.shape
        tg = tw = None
        out = None
        if taps:
            g = mbb_feed_geometry(self.roi, self.bands, tl, (iw, ih))
            W, H = g[2], g[3]
            gs, ws = [], []
            for _ in range(self.bands + 1):
                gs.append(np.empty((H, W, 3), np.int16))
                ws.append(np.empty((H, W), np.float32))
                W, H = (W + 1) // 2, (H + 1) // 2
            tg = (C.c_void_p * 16)(*[a.ctypes.data for a in gs])
            tw = (C.c_void_p * 16)(*[a.ctypes.data for a in ws])
            out = (g, gs, ws)
        lib().orc_mbb_feed(self._h, _p(img16), _p(mask), C.c_int(tl[0]), C.c_int(tl[1]), C.c_int(iw), C.c_int(ih),
                           tg, tw)
        return out

    def blend(self, taps=False):
        w, h = self.final
        out = np.empty((h, w, 3), np.int16)
        m = np.empty((h, w), np.uint8)
        tn = None
        norm = None
        if taps:
            norm = []
            W, H = self.roi[2], self.roi[3]
            for _ in range(self.bands + 1):
                norm.append(np.empty((H, W, 3), np.int16))
                W, H = (W + 1) // 2, (H + 1) // 2
            tn = (C.c_void_p * 16)(*[a.ctypes.data for a in norm])
        lib().orc_mbb_blend(self._h, _p(out), _p(m), tn)
        return (out, m, norm) if taps else (out, m)

    def __del__(self):
        try:
            lib().orc_mbb_destroy(self._h)
        except Exception:
            pass


class Feather:
    """cv::detail::FeatherBlender(sharpness) restated."""

    def __init__(self, roi, sharpness=0.02):
        x, y, w, h = roi
        self.roi = roi
        self._h = C.c_void_p(lib().orc_feather_create(C.c_int(x), C.c_int(y), C.c_int(w), C.c_int(h),
                                                       C.c_float(sharpness)))

    def feed(self, img16, mask, tl):
        img16 = np.ascontiguousarray(img16, np.int16)
        mask = np.ascontiguousarray(mask, np.uint8)
        ih, iw = mask.shape
        lib().orc_feather_feed(self._h, _p(img16), _p(mask), C.c_int(tl[0]), C.c_int(tl[1]), C.c_int(iw), C.c_int(ih),
                               None)

    def blend(self):
        w, h = self.roi[2], self.roi[3]
        out = np.empty((h, w, 3), np.int16)
        m = np.empty((h, w), np.uint8)
        lib().orc_feather_blend(self._h, _p(out), _p(m))
        return out, m

    def __del__(self):
        try:
            lib().orc_feather_destroy(self._h)
        except Exception:
            pass


def s16_to_u8(a):
    a = np.ascontiguousarray(a, np.int16)
    out = np.empty(a.shape, np.uint8)
    lib().orc_s16_to_u8(_p(a), C.c_size_t(a.size), _p(out))
    return out


def result_roi(corners, sizes):
    """cv::detail::resultRoi(corners, sizes) -> (x, y, w, h)."""
    tlx = min(c[0] for c in corners)
    tly = min(c[1] for c in corners)
    brx = max(c[0] + s[0] for c, s in zip(corners, sizes))
    bry = max(c[1] + s[1] for c, s in zip(corners, sizes))
    return (tlx, tly, brx - tlx, bry - tly)


def warp_frame(img, K, R, scale, affine=True):
    """warper.warp(img, LINEAR, REFLECT) + warper.warp(255s, NEAREST, CONSTANT).
    -> dict(corner, size, xy, a, warped, mask)"""
    h, w = img.shape[:2]
    k_rinv, r_kinv, t = projector_setup(K, R, affine)
    tlx, tly, brx, bry = plane_roi(r_kinv, t, scale, w, h)
    mw, mh = brx - tlx + 1, bry - tly + 1
    xm, ym = plane_maps(k_rinv, t, scale, tlx, tly, mw, mh)
    xy, a = fixed_tables(xm, ym)
    warped = remap_bilinear(np.ascontiguousarray(img), xy, a, "reflect")
    mask = nearest_mask(xm, ym, w, h)
    return dict(corner=(tlx, tly), size=(mw, mh), xy=xy, a=a, warped=warped, mask=mask)


def compose_port(frames, Ks, Rs, scale, blend="multiband", bands=5, sharpness=0.02, affine=True, taps=None):
    """The restated composePanorama hot loop. frames: list of HxWx3 uint8 BGR.
    Returns (pano_u8 HxWx3, mask HxW, roi(x,y,w,h))."""
    warps = [warp_frame(f, K, R, scale, affine) for f, K, R in zip(frames, Ks, Rs)]
    corners = [w["corner"] for w in warps]
    sizes = [w["size"] for w in warps]
    roi = result_roi(corners, sizes)
    bl = MultiBand(roi, bands) if blend == "multiband" else Feather(roi, sharpness)
    for w in warps:
        bl.feed(w["warped"].astype(np.int16), w["mask"], w["corner"])
    out16, m = bl.blend()
    if taps is not None:
        taps["warps"] = warps
    return s16_to_u8(out16), m, roi
