"""Host-side mirror of the reference's compose step over the C ABI.

`compose_panorama` takes what the reference has in hand right after
`stitcher->estimateTransform(images)` (/root/reference/src/stitch_robust.cpp:251) — the frames, the
per-camera K / R (float32 3x3) and the warp scale — and does what `stitcher->composePanorama(output)`
(:256) does with the components the reference configures (:203-213): AffineWarper placement,
LINEAR/REFLECT image warp, NEAREST/CONSTANT mask warp, MultiBandBlender(bands) (or
FeatherBlender(0.02) for BASELINE config 1), ->8U. All pixel work happens in libdronestitch_cuda.
"""
import ctypes as C

import numpy as np

from . import _lib as L


def plane_transform(K, R, scale, affine=True, border=L.DS_BORDER_REFLECT):
    t = L.ds_transform()
    t.kind = L.DS_XF_PLANE_F32
    t.affine_warper = 1 if affine else 0
    K = np.ascontiguousarray(K, np.float32).reshape(9)
    R = np.ascontiguousarray(R, np.float32).reshape(9)
    for i in range(9):
        t.K[i] = float(K[i])
        t.R[i] = float(R[i])
    t.scale = float(np.float32(scale))
    t.border = border
    return t


def affine_transform(M, corner, size, border=L.DS_BORDER_CONSTANT):
    """cv::warpAffine semantics: forward 2x3 double M (source px -> px of the frame's own bbox)."""
    t = L.ds_transform()
    t.kind = L.DS_XF_AFFINE_F64
    M = np.ascontiguousarray(M, np.float64).reshape(-1)
    for i in range(6):
        t.M[i] = float(M[i])
    t.M[6], t.M[7], t.M[8] = 0.0, 0.0, 1.0
    t.corner_x, t.corner_y = int(corner[0]), int(corner[1])
    t.width, t.height = int(size[0]), int(size[1])
    t.border = border
    return t


def homography_transform(H, corner, size, border=L.DS_BORDER_CONSTANT):
    """cv::warpPerspective semantics: forward 3x3 double H."""
    t = L.ds_transform()
    t.kind = L.DS_XF_HOMOGRAPHY_F64
    H = np.ascontiguousarray(H, np.float64).reshape(9)
    for i in range(9):
        t.M[i] = float(H[i])
    t.corner_x, t.corner_y = int(corner[0]), int(corner[1])
    t.width, t.height = int(size[0]), int(size[1])
    t.border = border
    return t


def warp_roi(xf, w, h, lib=None):
    lib = lib or L.default_library()
    out = (C.c_int32 * 4)()
    lib.check(lib.dll.ds_warp_roi(C.byref(xf), int(w), int(h), out))
    return tuple(out)


def result_roi(rois):
    """cv::detail::resultRoi(corners, sizes) over (x, y, w, h) tuples."""
    x0 = min(r[0] for r in rois)
    y0 = min(r[1] for r in rois)
    x1 = max(r[0] + r[2] for r in rois)
    y1 = max(r[1] + r[3] for r in rois)
    return (x0, y0, x1 - x0, y1 - y0)


def global_blend_bands(canvas_w, canvas_h, configured_bands, lib=None):
    """Band count of the global stage's blender (/root/reference/src/stitch_global.cpp:632-635)."""
    lib = lib or L.default_library()
    return int(lib.dll.ds_global_blend_bands(int(canvas_w), int(canvas_h), int(configured_bands)))


def plan_row_bands(roi, rois, n_bands, blend="multiband", bands=5, lib=None):
    """Row-band edges (ds_plan_row_bands) for `n_bands` handles of one canvas, balanced by the frames' footprints."""
    lib = lib or L.default_library()
    d = L.ds_canvas_desc()
    d.x, d.y, d.width, d.height = [int(v) for v in roi]
    d.blend_mode = L.DS_BLEND_MULTIBAND if blend == "multiband" else L.DS_BLEND_FEATHER
    d.num_bands = int(bands)
    d.sharpness = 0.02
    flat = (C.c_int32 * (4 * max(len(rois), 1)))(*[int(v) for r in rois for v in r])
    out = (C.c_int32 * (n_bands + 1))()
    lib.check(lib.dll.ds_plan_row_bands(C.byref(d), flat, len(rois), int(n_bands), out))
    return list(out)


def warp_frame(img, xf, device=0, lib=None):
    """ds_warp_frame: one frame warped on its own (the seam-phase warps of composePanorama, a strip's warpAffine).
    -> (corner (x, y), warped HxWx3 uint8, warped mask HxW uint8)."""
    lib = lib or L.default_library()
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3 and img.strides[2] == 1 and img.strides[1] == 3
    pl = (C.c_int32 * 4)()
    args = (int(device), C.c_void_p(img.ctypes.data), img.shape[1], img.shape[0], img.strides[0], C.byref(xf), pl)
    lib.check(lib.dll.ds_warp_frame(*args, None, None))
    out = np.empty((pl[3], pl[2], 3), np.uint8)
    mask = np.empty((pl[3], pl[2]), np.uint8)
    lib.check(lib.dll.ds_warp_frame(*args, out.ctypes.data, mask.ctypes.data))
    return (pl[0], pl[1]), out, mask


class Canvas:
    """One ds_canvas handle (one GPU, one row band)."""

    def __init__(self, roi, blend="multiband", bands=5, sharpness=0.02, out_format="bgr", device=0, band=None,
                 stream=None, lib=None, pipeline_rows=0):
        self.lib = lib or L.default_library()
        d = L.ds_canvas_desc()
        d.x, d.y, d.width, d.height = [int(v) for v in roi]
        d.blend_mode = L.DS_BLEND_MULTIBAND if blend == "multiband" else L.DS_BLEND_FEATHER
        d.num_bands = int(bands)
        d.sharpness = float(sharpness)
        d.out_format = L.DS_OUT_BGRA8 if out_format == "bgra" else L.DS_OUT_BGR8
        d.device = int(device)
        if band is not None:
            d.band_y0, d.band_y1 = int(band[0]), int(band[1])
        d.stream = C.c_void_p(stream) if stream else None
        d.pipeline_rows = int(pipeline_rows)   # 0 auto, > 0 always row slices of about that height, < 0 never
        self.desc = d
        self._pending = []   # buffers lent to asynchronous uploads
        self.roi = tuple(int(v) for v in roi)
        self.bpp = 4 if out_format == "bgra" else 3
        self._h = C.c_void_p()
        self.lib.check(self.lib.dll.ds_create_canvas(C.byref(d), C.byref(self._h)))

    def close(self):
        if self._h:
            self.lib.dll.ds_destroy_canvas(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def touches(self, frame_roi):
        arr = (C.c_int32 * 4)(*[int(v) for v in frame_roi])
        return bool(self.lib.dll.ds_frame_touches_band(C.byref(self.desc), arr))

    @staticmethod
    def _make_opts(seam_mask=None, channel_gain=None, seam_lowres=None, compensator_gain=None, gain_map=None, async_=False,
                   content_mask=False, seam_nearest=False, soft_mask=None, gain_blocks=None):
        """-> (ds_frame_opts or None, buffers to keep alive)."""
        keep = []
        if not (async_ or content_mask or seam_nearest or soft_mask is not None or
                any(v is not None for v in (seam_mask, channel_gain, seam_lowres, compensator_gain, gain_map, gain_blocks))):
            return None, keep
        opts = L.ds_frame_opts()
        opts.flags = ((L.DS_UPLOAD_ASYNC if async_ else 0) | (L.DS_MASK_CONTENT if content_mask else 0) |
                      (L.DS_SEAM_NEAREST if seam_nearest else 0) | (L.DS_MASK_SOFT if soft_mask is not None else 0))
        if soft_mask is not None and soft_mask is not True:
            opts.soft_sigma = float(soft_mask)
        if compensator_gain is not None:
            cg = (C.c_double * 3)(*[float(v) for v in compensator_gain])
            keep.append(cg)
            opts.compensator_gain = C.cast(cg, C.POINTER(C.c_double))
        if gain_map is not None:
            gm = np.ascontiguousarray(gain_map, np.float32)
            keep.append(gm)
            opts.gain_map = gm.ctypes.data
            opts.gain_map_stride = gm.strides[0]
        if gain_blocks is not None:
            gb = np.ascontiguousarray(gain_blocks, np.float32)
            keep.append(gb)
            opts.gain_blocks = gb.ctypes.data
            opts.gain_blocks_w, opts.gain_blocks_h = gb.shape[1], gb.shape[0]
            opts.gain_blocks_stride = gb.strides[0]
        if seam_lowres is not None:
            sl = np.ascontiguousarray(seam_lowres, np.uint8)
            keep.append(sl)
            opts.seam_lowres = sl.ctypes.data
            opts.seam_lowres_w, opts.seam_lowres_h = sl.shape[1], sl.shape[0]
            opts.seam_lowres_stride = sl.strides[0]
        if seam_mask is not None:
            sm = np.ascontiguousarray(seam_mask, np.uint8)
            keep.append(sm)
            opts.seam_mask = sm.ctypes.data
            opts.seam_mask_stride = sm.strides[0]
        if channel_gain is not None:
            g = (C.c_float * 3)(*[float(v) for v in channel_gain])
            keep.append(g)
            opts.channel_gain = C.cast(g, C.POINTER(C.c_float))
        return opts, keep

    def upload(self, idx, img, xf, async_=False, **opt_kw):
        """img: HxWx3 uint8 (numpy; any row stride) or a (ptr, w, h, stride) tuple.
        async_: DS_UPLOAD_ASYNC - `img` (pinned) must stay valid and unchanged until composite() / synchronize() /
        a full download() has returned.
        opt_kw (ds_frame_opts): seam_mask, seam_lowres, seam_nearest, channel_gain, compensator_gain, gain_map, gain_blocks,
        content_mask (DS_MASK_CONTENT), soft_mask (DS_MASK_SOFT: True or the sigma)."""
        opts, keep = self._make_opts(async_=async_, **opt_kw)
        if isinstance(img, tuple):
            ptr, w, h, stride = img
        else:
            assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3 and img.strides[2] == 1 and img.strides[1] == 3
            ptr, w, h, stride = img.ctypes.data, img.shape[1], img.shape[0], img.strides[0]
        self.lib.check(self.lib.dll.ds_upload_frame(self._h, int(idx), C.c_void_p(ptr), int(w), int(h), int(stride),
                                                    C.byref(xf), C.byref(opts) if opts is not None else None))
        if async_:
            self._pending.append((img, keep))

    def update_opts(self, idx, **opt_kw):
        """ds_update_frame_opts: new seam masks / gains / mask flags for a frame whose pixels are already uploaded
        (the global stage's feed loop, /root/reference/src/stitch_global.cpp:643-660)."""
        opts, keep = self._make_opts(**opt_kw)
        self.lib.check(self.lib.dll.ds_update_frame_opts(self._h, int(idx), C.byref(opts) if opts is not None else None))

    def frame_mask(self, idx, which=0):
        """ds_download_frame_mask: 0 = the mask the blender is fed with, 1 = the content mask (DS_MASK_CONTENT)."""
        x, y, w, h = self.placement(idx)
        m = np.empty((h, w), np.uint8)
        self.lib.check(self.lib.dll.ds_download_frame_mask(self._h, int(idx), int(which), m.ctypes.data, m.strides[0]))
        return m

    def upload_device(self, idx, dev_ptr, w, h, stride, xf):
        self.lib.check(self.lib.dll.ds_upload_frame_device(self._h, int(idx), C.c_void_p(dev_ptr), int(w), int(h),
                                                           int(stride), C.byref(xf), None))

    def composite(self):
        self.lib.check(self.lib.dll.ds_composite(self._h))
        self._pending.clear()

    def composite_async(self):
        self.lib.check(self.lib.dll.ds_composite_async(self._h))

    # ---- NVLink P2P halo exchange between row-band handles (include/dronestitch.h)
    def p2p_export(self):
        """-> bytes describing this handle's band, counters and per-frame level-1 arrays, for the neighbours."""
        n = C.c_size_t()
        self.lib.check(self.lib.dll.ds_p2p_export(self._h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        self.lib.check(self.lib.dll.ds_p2p_export(self._h, buf, n.value, C.byref(n)))
        return buf.raw[:n.value]

    def p2p_connect(self, side, blob):
        """side 0: the handle of the band above, 1: below."""
        self.lib.check(self.lib.dll.ds_p2p_connect(self._h, int(side), blob, len(blob)))

    def p2p_disconnect(self):
        self.lib.check(self.lib.dll.ds_p2p_disconnect(self._h))

    def composite_stage(self, stage):
        self.lib.check(self.lib.dll.ds_composite_stage(self._h, int(stage)))

    def synchronize(self):
        self.lib.check(self.lib.dll.ds_synchronize(self._h))
        self._pending.clear()

    def info(self):
        i = L.ds_canvas_info()
        self.lib.check(self.lib.dll.ds_get_info(self._h, C.byref(i)))
        return i

    def download(self, x=0, y=None, w=None, h=None, want_mask=True, out=None, mask_out=None):
        i = self.info()
        y0 = i.band_y0 if y is None else y
        y1 = min(i.band_y1, self.roi[3])
        w = self.roi[2] - x if w is None else w
        h = y1 - y0 if h is None else h
        if out is None:
            out = np.empty((h, w, self.bpp), np.uint8)
        mask = mask_out
        if mask is None and want_mask and self.bpp == 3:
            mask = np.empty((h, w), np.uint8)
        self.lib.check(self.lib.dll.ds_download_tile(self._h, int(x), int(y0), int(w), int(h), out.ctypes.data,
                                                     out.strides[0], mask.ctypes.data if mask is not None else None,
                                                     mask.strides[0] if mask is not None else 0))
        return out, mask

    def auto_crop_rect(self):
        """ds_auto_crop_rect: the rectangle autoCropBlackBorder (/root/reference/src/stitch_common.cpp:4-27) keeps,
        (x, y, w, h) relative to the canvas origin; download it with download(x, y, w, h)."""
        o = (C.c_int32 * 4)()
        self.lib.check(self.lib.dll.ds_auto_crop_rect(self._h, o))
        return tuple(o)

    def set_profiling(self, on=True):
        self.lib.check(self.lib.dll.ds_set_profiling(self._h, 1 if on else 0))

    def kernel_times(self):
        """-> list of dict(name, level, ms, algorithmic_bytes) for the last profiled composite."""
        n = C.c_int()
        arr = (L.ds_kernel_time * 8192)()
        self.lib.check(self.lib.dll.ds_get_kernel_times(self._h, arr, 8192, C.byref(n)))
        return [dict(name=arr[i].name.decode(), level=arr[i].level, ms=arr[i].ms, algorithmic_bytes=arr[i].algorithmic_bytes)
                for i in range(min(n.value, 8192))]

    # ---- debug taps
    def placement(self, idx):
        o = (C.c_int32 * 4)()
        self.lib.check(self.lib.dll.ds_debug_get_placement(self._h, int(idx), o))
        return tuple(o)

    def maps(self, idx):
        x, y, w, h = self.placement(idx)
        xy = np.empty((h, w, 2), np.int16)
        a = np.empty((h, w), np.uint16)
        self.lib.check(self.lib.dll.ds_debug_get_maps(self._h, int(idx), xy.ctypes.data, a.ctypes.data))
        return xy, a

    def warped(self, idx):
        x, y, w, h = self.placement(idx)
        img = np.empty((h, w, 3), np.uint8)
        m = np.empty((h, w), np.uint8)
        self.lib.check(self.lib.dll.ds_debug_get_warped(self._h, int(idx), img.ctypes.data, m.ctypes.data))
        return img, m

    def frame_level(self, idx, level):
        dims = (C.c_int32 * 4)()
        self.lib.check(self.lib.dll.ds_debug_get_frame_level(self._h, int(idx), int(level), None, None, dims))
        g = np.empty((dims[3], dims[2], 3), np.int16)
        w = np.empty((dims[3], dims[2]), np.float32)
        self.lib.check(self.lib.dll.ds_debug_get_frame_level(self._h, int(idx), int(level), g.ctypes.data, w.ctypes.data, dims))
        return g, w, tuple(dims)


def compose_panorama(images, Ks, Rs, scale, blend="multiband", bands=5, sharpness=0.02, affine=True,
                     out_format="bgr", device=0, lib=None, return_canvas=False, seam_lowres=None):
    """The compose half of stitchWithMode (/root/reference/src/stitch_robust.cpp:255-256).
    seam_lowres: optional list of the low-resolution seam masks cv::Stitcher holds after seam finding (one per image;
    dilated, resized with INTER_LINEAR_EXACT and ANDed into the warped mask on the device, as composePanorama does).
    Returns (pano HxWx3 uint8, result_mask HxW uint8, roi (x, y, w, h))."""
    lib = lib or L.default_library()
    xfs = [plane_transform(K, R, scale, affine) for K, R in zip(Ks, Rs)]
    rois = [warp_roi(xf, im.shape[1], im.shape[0], lib) for xf, im in zip(xfs, images)]
    roi = result_roi(rois)
    cv = Canvas(roi, blend, bands, sharpness, out_format, device, lib=lib)
    # the images outlive this call: queue the uploads, let the composite chase them slice by slice, and copy each
    # slice out as it completes (DS_UPLOAD_ASYNC, include/dronestitch.h)
    for i, (im, xf) in enumerate(zip(images, xfs)):
        if seam_lowres is not None:
            cv.upload(i, im, xf, async_=True, seam_lowres=seam_lowres[i])
        else:
            cv.upload(i, im, xf, async_=True)
    cv.composite_async()
    pano, mask = cv.download()
    cv.synchronize()
    if return_canvas:
        return pano, mask, roi, cv
    cv.close()
    return pano, mask, roi


def compose_panorama_from_stitcher(images, cameras, work_scale, warped_image_scale=None, blend="multiband", bands=5,
                                   affine=True, seam_lowres=None, lib=None, device=0):
    """Python twin of ds::composePanorama (include/dronestitch.hpp): what replaces `stitcher->composePanorama(output)`
    at /root/reference/src/stitch_robust.cpp:256. `cameras` and `work_scale` are what cv::Stitcher holds after
    estimateTransform (:251) - objects with focal / aspect / ppx / ppy / R like cv::detail::CameraParams;
    warped_image_scale defaults to the median focal, as cv::Stitcher computes it. Full-resolution compositing
    (compositing_resol_mpx = -1): the cameras are rescaled by compose_work_aspect = 1 / work_scale."""
    focals = sorted(float(c.focal) for c in cameras)
    if warped_image_scale is None:
        n = len(focals)
        warped_image_scale = focals[n // 2] if n % 2 == 1 else (focals[n // 2 - 1] + focals[n // 2]) * 0.5
    cwa = 1.0 / float(work_scale)
    scale = np.float32(float(np.float32(warped_image_scale)) * cwa)
    Ks, Rs = [], []
    for c in cameras:
        focal, ppx, ppy = float(c.focal) * cwa, float(c.ppx) * cwa, float(c.ppy) * cwa
        Ks.append(np.array([[focal, 0, ppx], [0, focal * float(c.aspect), ppy], [0, 0, 1]], np.float64).astype(np.float32))
        Rs.append(np.ascontiguousarray(c.R, np.float32))
    return compose_panorama(images, Ks, Rs, scale, blend, bands, affine=affine, lib=lib, device=device, seam_lowres=seam_lowres)
