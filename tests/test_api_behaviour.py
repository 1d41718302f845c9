"""Handle life-cycle behaviour through the C ABI: re-upload, sparse frame indices, repeated composites,
tile download, per-kernel profile. Run on the emulator (host logic is shared) and, marked gpu, on the GPU
(plus the device-pointer upload, which only exists there)."""
import numpy as np
import pytest

from drone_image_stitch_cpp_b200 import compositor as CP
from drone_image_stitch_cpp_b200 import synth
from oracle import ds_oracle as O


def _behaviour(lib):
    sv = synth.grid_survey(2, 2, 240, 180, overlap=0.6, seed=41, work_scale=0.5)
    xfs = [CP.plane_transform(K, R, sv.scale) for K, R in zip(sv.Ks, sv.Rs)]
    rois = [CP.warp_roi(xf, 240, 180, lib) for xf in xfs]
    roi = CP.result_roi(rois)
    ref, refmask, _ = O.compose_port(sv.frames, sv.Ks, sv.Rs, sv.scale, "multiband", 4)
    cv = CP.Canvas(roi, "multiband", 4, lib=lib)
    # sparse indices (feed order = index order), uploaded out of order
    idx = [7, 2, 11, 30]
    order = sorted(range(4), key=lambda i: idx[i])          # frame i gets index idx[i]; feed order follows indices
    for i in (3, 0, 2, 1):
        cv.upload(idx[i], sv.frames[i], xfs[i])
    cv.composite()
    pano, mask = cv.download()
    ref2, refmask2, _ = O.compose_port([sv.frames[i] for i in order], [sv.Ks[i] for i in order], [sv.Rs[i] for i in order],
                                       sv.scale, "multiband", 4)
    assert np.array_equal(pano, ref2) and np.array_equal(mask, refmask2)
    # repeated composite is idempotent
    cv.composite()
    pano_b, _ = cv.download()
    assert np.array_equal(pano_b, pano)
    # re-upload an index with different pixels: result changes accordingly
    other = np.ascontiguousarray(sv.frames[1][::-1, ::-1])
    cv.upload(idx[1], other, xfs[1])
    cv.composite()
    pano_c, _ = cv.download()
    fr = list(sv.frames)
    fr[1] = other
    ref3, _, _ = O.compose_port([fr[i] for i in order], [sv.Ks[i] for i in order], [sv.Rs[i] for i in order], sv.scale, "multiband", 4)
    assert np.array_equal(pano_c, ref3)
    # tile download equals the crop of the full download
    t, tm = cv.download(x=33, y=21, w=100, h=77)
    assert np.array_equal(t, pano_c[21:98, 33:133])
    # strided host buffers (cv::Mat::step > 3*w) are accepted
    wide = np.zeros((180, 300, 3), np.uint8)
    wide[:, :240] = sv.frames[0]
    cv.upload(idx[0], wide[:, :240], xfs[0])
    # per-kernel profile covers every launch of a composite
    cv.set_profiling(True)
    cv.composite()
    kt = cv.kernel_times()
    assert len(kt) == cv.info().launches_last_composite and any(k["name"] == "mb_feed" for k in kt)
    cv.close()
    return ref, refmask


def test_behaviour_emu(emu_lib):
    _behaviour(emu_lib)


@pytest.mark.gpu
def test_behaviour_gpu(cuda_lib):
    _behaviour(cuda_lib)


@pytest.mark.gpu
def test_upload_from_device_pointer(cuda_lib):
    import torch
    sv = synth.grid_survey(2, 1, 320, 240, overlap=0.5, seed=42)
    xfs = [CP.plane_transform(K, R, sv.scale) for K, R in zip(sv.Ks, sv.Rs)]
    roi = CP.result_roi([CP.warp_roi(xf, 320, 240, cuda_lib) for xf in xfs])
    ref, refmask, _ = O.compose_port(sv.frames, sv.Ks, sv.Rs, sv.scale, "multiband", 3)
    cv = CP.Canvas(roi, "multiband", 3, lib=cuda_lib)
    keep = []
    for i, f in enumerate(sv.frames):
        t = torch.from_numpy(f).cuda()
        keep.append(t)
        torch.cuda.synchronize()
        cv.upload_device(i, t.data_ptr(), 320, 240, 320 * 3, xfs[i])
    cv.composite()
    pano, mask = cv.download()
    assert np.array_equal(pano, ref) and np.array_equal(mask, refmask)
    cv.close()


def _edges(lib):
    # empty input: a canvas composited with no frame is all zero with an all-zero mask (what blend() gives
    # when nothing was fed)
    for blend in ("multiband", "feather"):
        cv = CP.Canvas((5, -3, 70, 45), blend, 3, lib=lib)
        cv.composite()
        pano, mask = cv.download()
        assert pano.shape == (45, 70, 3) and not pano.any() and not mask.any()
        cv.close()
    # tiny ragged frames: 9x7 and 1x1 sources, canvas smaller than one tile, bands cropped by the canvas size
    rng = np.random.default_rng(5)
    frames = [rng.integers(0, 256, (7, 9, 3)).astype(np.uint8), rng.integers(0, 256, (1, 1, 3)).astype(np.uint8)]
    Ks = [np.eye(3, dtype=np.float32)] * 2
    Rs = [np.array([[1, 0, 2.25], [0, 1, -1.5], [0, 0, 1]], np.float32), np.array([[1, 0, 6.0], [0, 1, 3.0], [0, 0, 1]], np.float32)]
    for blend, bands in (("multiband", 5), ("feather", 0)):
        ref, refmask, roi = O.compose_port(frames, Ks, Rs, 1.0, blend, bands)
        pano, mask, roi2 = CP.compose_panorama(frames, Ks, Rs, 1.0, blend, bands, lib=lib)
        assert roi == roi2 and np.array_equal(mask, refmask) and np.array_equal(pano, ref)
    # maximum source size: cv::remap asserts < 32768 per side; the library refuses instead of asserting
    cv = CP.Canvas((0, 0, 64, 64), "multiband", 2, lib=lib)
    from drone_image_stitch_cpp_b200 import _lib as L
    with pytest.raises(L.DroneStitchError):
        cv.upload(0, (0x1000, 40000, 2, 40000 * 3), CP.plane_transform(np.eye(3), np.eye(3), 1.0))
    cv.close()


def test_edges_emu(emu_lib):
    _edges(emu_lib)


@pytest.mark.gpu
def test_edges_gpu(cuda_lib):
    _edges(cuda_lib)


def _pipelined(lib, blend):
    """DS_UPLOAD_ASYNC + ds_composite_async + ds_download_tile: the sliced schedule, tile downloads that cross slice
    edges, a second step into the same handle (same geometry: launch metadata is reused), and the switch back to
    synchronous uploads - always the bytes of the plain schedule."""
    sv = synth.grid_survey(2, 3, 200, 160, overlap=0.5, seed=43, work_scale=0.5)
    xfs = [CP.plane_transform(K, R, sv.scale) for K, R in zip(sv.Ks, sv.Rs)]
    rois = [CP.warp_roi(xf, 200, 160, lib) for xf in xfs]
    roi = CP.result_roi(rois)
    ref, refmask, _ = O.compose_port(sv.frames, sv.Ks, sv.Rs, sv.scale, blend, 3)
    other = [np.ascontiguousarray(f[::-1]) for f in sv.frames]
    ref_b, refmask_b, _ = O.compose_port(other, sv.Ks, sv.Rs, sv.scale, blend, 3)
    for rows in (32, 96):
        cv = CP.Canvas(roi, blend, 3, lib=lib, pipeline_rows=rows)
        for i, (f, xf) in enumerate(zip(sv.frames, xfs)):
            cv.upload(i, f, xf, async_=True)
        cv.composite_async()
        # a tile that starts and ends inside slices
        y0, h = roi[3] // 3 + 5, roi[3] // 2
        x0, w = 7, roi[2] - 20
        tile, tmask = cv.download(x=x0, y=y0, w=w, h=h)
        assert np.array_equal(tile, ref[y0:y0 + h, x0:x0 + w]) and np.array_equal(tmask, refmask[y0:y0 + h, x0:x0 + w])
        pano, mask = cv.download()
        cv.synchronize()
        assert np.array_equal(pano, ref) and np.array_equal(mask, refmask)
        n_sliced = cv.info().launches_last_composite
        # next step: new pixels, same slots and geometry
        for i, (f, xf) in enumerate(zip(other, xfs)):
            cv.upload(i, f, xf, async_=True)
        cv.composite_async()
        pano, mask = cv.download()
        cv.synchronize()
        assert np.array_equal(pano, ref_b) and np.array_equal(mask, refmask_b)
        assert cv.info().launches_last_composite == n_sliced
        cv.close()
    # automatic mode: slices only while asynchronous uploads are pending; never with pipeline_rows < 0
    cv = CP.Canvas(roi, blend, 3, lib=lib, pipeline_rows=-1)
    for i, (f, xf) in enumerate(zip(sv.frames, xfs)):
        cv.upload(i, f, xf, async_=True)
    cv.composite()
    pano, mask = cv.download()
    assert np.array_equal(pano, ref) and np.array_equal(mask, refmask)
    assert cv.info().launches_last_composite < n_sliced
    cv.close()


@pytest.mark.parametrize("blend", ["multiband", "feather"])
def test_pipelined_emu(emu_lib, blend):
    _pipelined(emu_lib, blend)


@pytest.mark.gpu
@pytest.mark.parametrize("blend", ["multiband", "feather"])
def test_pipelined_gpu(cuda_lib, blend):
    _pipelined(cuda_lib, blend)


def _band_partial_upload(lib):
    """A row-band handle fed with DS_UPLOAD_ASYNC copies only the source rows that map into its band (+ halo):
    fewer bytes than the frames hold, the same output rows, and the whole-frame debug tap refuses."""
    import os
    from drone_image_stitch_cpp_b200 import _lib as L
    sv = synth.grid_survey(1, 4, 160, 240, overlap=0.3, seed=47, work_scale=0.5)
    xfs = [CP.plane_transform(K, R, sv.scale) for K, R in zip(sv.Ks, sv.Rs)]
    rois = [CP.warp_roi(xf, 160, 240, lib) for xf in xfs]
    roi = CP.result_roi(rois)
    whole = CP.Canvas(roi, "multiband", 2, lib=lib)
    for i, (f, xf) in enumerate(zip(sv.frames, xfs)):
        whole.upload(i, f, xf)
    whole.composite()
    ref, _ = whole.download()
    assert whole.info().h2d_bytes_total == sum(f.nbytes for f in sv.frames)
    H = whole.info().padded_height
    y0, y1 = (H // 3) // 4 * 4, (2 * H // 3) // 4 * 4
    os.environ["DS_UPLOAD_CHUNK_ROWS"] = "16"
    try:
        band = CP.Canvas(roi, "multiband", 2, band=(y0, y1), lib=lib)
    finally:
        os.environ.pop("DS_UPLOAD_CHUNK_ROWS", None)
    sent = 0
    for i, (f, xf) in enumerate(zip(sv.frames, xfs)):
        if band.touches(rois[i]):
            band.upload(i, f, xf, async_=True)
            sent += f.nbytes
    band.composite_async()
    rows, _ = band.download()
    band.synchronize()
    assert np.array_equal(rows, ref[y0:min(y1, roi[3])])
    copied = band.info().h2d_bytes_total
    assert 0 < copied < 0.8 * sent, (copied, sent)
    # a second composite over the resident rows gives the same bytes
    band.composite()
    rows2, _ = band.download()
    assert np.array_equal(rows2, rows)
    with pytest.raises(L.DroneStitchError):
        band.warped(0 if band.touches(rois[0]) else 1)
    band.close()
    whole.close()


def test_band_partial_upload_emu(emu_lib):
    _band_partial_upload(emu_lib)


@pytest.mark.gpu
def test_band_partial_upload_gpu(cuda_lib):
    _band_partial_upload(cuda_lib)


def _widened_api_errors(lib):
    """Error behaviour of the entry points of SURVEY 8(f): states and arguments are refused with a status, never guessed."""
    from drone_image_stitch_cpp_b200 import _lib as L
    import ctypes as C
    img = synth.orthophoto(120, 160, 3).numpy()
    M = np.array([[1.0, 0.02, 3.5], [-0.02, 1.0, 2.25]])
    xf = CP.affine_transform(M, (0, 0), (170, 130))
    cv = CP.Canvas((0, 0, 170, 130), "multiband", 3, lib=lib)

    def expect(code, fn):
        with pytest.raises(L.DroneStitchError) as e:
            fn()
        assert e.value.code == code, e.value

    expect(L.DS_ERR_BAD_ARG, lambda: cv.update_opts(0, content_mask=True))        # nothing uploaded at index 0
    cv.upload(0, img, xf)
    expect(L.DS_ERR_STATE, lambda: cv.frame_mask(0, 1))                           # no DS_MASK_CONTENT on that frame
    expect(L.DS_ERR_BAD_ARG, lambda: cv.frame_mask(0, 2))
    expect(L.DS_ERR_UNSUPPORTED, lambda: cv.update_opts(0, soft_mask=12.0))       # kernel wider than 81 taps
    expect(L.DS_ERR_STATE, lambda: cv.auto_crop_rect())                           # before any composite
    # a frame without options: mask 0 is the nearest-warped 255s
    m = cv.frame_mask(0, 0)
    assert np.array_equal(m, O.affine_nearest_mask(M, 170, 130, 160, 120))
    # a smaller sigma is accepted and matches the oracle
    cv.update_opts(0, soft_mask=3.0, content_mask=True)
    assert np.array_equal(cv.frame_mask(0, 0), O.soft_blend_mask(None, O.content_mask(img, M, 170, 130), 3.0))
    cv.composite()
    assert cv.auto_crop_rect() == O.auto_crop_rect(cv.download()[0])
    # a black canvas is returned whole, like the reference leaves the panorama untouched
    cv.upload(0, np.zeros_like(img), xf)
    cv.composite()
    assert cv.auto_crop_rect() == (0, 0, 170, 130)
    cv.close()
    # ds_warp_frame: placement query, argument checks
    pl = (C.c_int32 * 4)()
    lib.check(lib.dll.ds_warp_frame(0, C.c_void_p(img.ctypes.data), 160, 120, img.strides[0], C.byref(xf), pl, None, None))
    assert tuple(pl) == (0, 0, 170, 130)
    out = np.empty((130, 170, 3), np.uint8)
    assert lib.dll.ds_warp_frame(0, C.c_void_p(img.ctypes.data), 160, 120, img.strides[0], C.byref(xf), pl, out.ctypes.data, None) == L.DS_ERR_BAD_ARG


def test_widened_api_errors_emu(emu_lib):
    _widened_api_errors(emu_lib)


@pytest.mark.gpu
def test_widened_api_errors_gpu(cuda_lib):
    _widened_api_errors(cuda_lib)


def _plan_rows(L, H, band):
    """Python twin of plan_rows (csrc/ds_runtime.cu): rows of every level a band accumulates (acc) and the rows of the
    per-frame planes it must hold (own); own[0] = the level-0 tile rows that run."""
    lh = [H]
    for _ in range(L):
        lh.append((lh[-1] + 1) // 2)
    clip = lambda a, b, n: (max(a, 0), min(b, n))
    acc = [band]
    for l in range(1, L + 1):
        acc.append(clip((acc[l - 1][0] >> 1) - 1, ((acc[l - 1][1] - 1) >> 1) + 2, lh[l]))
    own = [None] * (L + 1)
    own[L] = acc[L]
    for l in range(L - 1, 0, -1):
        own[l] = clip(min(acc[l][0], 2 * own[l + 1][0] - 2), max(acc[l][1], 2 * own[l + 1][1] + 1), lh[l])
    own[0] = clip(min(band[0], 2 * own[1][0]) & ~1, (max(band[1], 2 * own[1][1]) + 1) & ~1, lh[0])
    return acc, own


def test_band_halo_rows(emu_lib):
    """Which frames a row band needs (ds_frame_touches_band). With every per-frame plane pixel computed once
    (ds_mb_pyrdown) the level-0 rows a band recomputes beyond its edges stay below 4 * 2^L per edge (124 / 94 rows at
    L = 5; the tile ring per level of round 1 needed ~290); a frame touches the band iff its feed ROI - bbox + the
    blender's 3 * 2^L gap, aligned to 2^L - reaches into those rows."""
    for L in (5, 8):
        m = 1 << L
        H = 64 * m
        y0, y1 = 24 * m, 40 * m
        _, own = _plan_rows(L, H, (y0, y1))
        lo, hi = own[0]
        assert 0 < y0 - lo <= 4 * m and 0 < hi - y1 <= 4 * m, (L, own[0])
        band = CP.Canvas((0, 0, 4096, H), "multiband", L, band=(y0, y1), lib=emu_lib)
        gap = 3 * m
        # a frame whose bbox is rows [0, h): its feed ROI ends at h + gap rounded up to 2^L
        e_out = (lo // m) * m                  # ROI end <= lo: outside
        assert not band.touches((100, 0, 500, e_out - gap)), (L, lo, e_out)
        assert band.touches((100, 0, 500, e_out - gap + m)), (L, lo, e_out)
        # a frame starting at row s: its feed ROI starts at (s - gap) rounded down to 2^L
        s_out = -(-hi // m) * m + gap          # ROI start >= hi: outside
        assert not band.touches((100, s_out, 500, 300)), (L, hi, s_out)
        assert band.touches((100, s_out - m, 500, 300)), (L, hi, s_out)
        band.close()
