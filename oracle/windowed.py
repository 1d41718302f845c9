"""TEST INFRASTRUCTURE (checker only): the windowed oracle of SURVEY.md 8(c) / probe P17 for canvases too large to blend
whole on the CPU. The blender (oracle/ds_oracle.c, the C restatement of cv::detail::MultiBandBlender, pinned against
cv2 in tests/test_oracle_vs_cv2.py) is prepared on a 2^bands-aligned window of the canvas, every warped frame is cropped
to the window, and the pixels at least 8 * 2^bands inside the window equal the full-canvas result bit for bit
(checked against the full-canvas oracle in tests/parity_cases.case_windowed_oracle).

Used by tests/ and by bench.py's in-run parity check; never by the product path.
Reference call order it follows: /root/reference/src/stitch_robust.cpp:256 (composePanorama's warp + feed loop with the
components of :203-213) and src/stitch_global.cpp:636-666 (prepare / feed / blend / ->8U)."""
import numpy as np

from . import ds_oracle as O


def margin(bands):
    return 8 << bands


def warp_into_window(img, K, R, scale, window_abs):
    """Frame warped (LINEAR / REFLECT image, NEAREST / CONSTANT mask) only over the part of its bbox inside `window_abs`
    = (x, y, w, h) in absolute canvas coordinates. -> (img16, mask, (x0, y0)) or None if it misses the window."""
    h, w = img.shape[:2]
    k_rinv, r_kinv, t = O.projector_setup(K, R, True)
    tlx, tly, brx, bry = O.plane_roi(r_kinv, t, scale, w, h)
    ax, ay, aw, ah = window_abs
    x0, y0, x1, y1 = max(tlx, ax), max(tly, ay), min(brx + 1, ax + aw), min(bry + 1, ay + ah)
    if x0 >= x1 or y0 >= y1:
        return None
    xm, ym = O.plane_maps(k_rinv, t, scale, x0, y0, x1 - x0, y1 - y0)
    xy, a = O.fixed_tables(xm, ym)
    warped = O.remap_bilinear(np.ascontiguousarray(img), xy, a, "reflect")
    mask = O.nearest_mask(xm, ym, w, h)
    return warped.astype(np.int16), mask, (x0, y0)


def compose_window(frames, Ks, Rs, scale, bands, roi, win):
    """frames: dict index -> HxWx3 uint8 (at least every frame that intersects the window), or a list.
    roi = canvas ROI (x, y, w, h), win = (x, y, w, h) relative to the canvas origin, all multiples of 2^bands.
    -> (pano u8, mask) of the window."""
    wx, wy, ww, wh = win
    m = 1 << bands
    assert wx % m == 0 and wy % m == 0 and ww % m == 0 and wh % m == 0
    window_abs = (roi[0] + wx, roi[1] + wy, ww, wh)
    bl = O.MultiBand(window_abs, bands)
    assert bl.bands == bands
    items = frames.items() if isinstance(frames, dict) else enumerate(frames)
    for i, f in sorted(items):
        got = warp_into_window(f, Ks[i], Rs[i], scale, window_abs)
        if got is not None:
            bl.feed(got[0], got[1], got[2])
    ref16, refmask = bl.blend()
    return O.s16_to_u8(ref16), refmask


def frames_touching(rois, roi, win):
    """Indices of the frames whose warped bbox intersects the window."""
    ax, ay, aw, ah = roi[0] + win[0], roi[1] + win[1], win[2], win[3]
    return [i for i, (x, y, w, h) in enumerate(rois) if x < ax + aw and x + w > ax and y < ay + ah and y + h > ay]


def compare_inside(pano_rows, mask_rows, rows_y0, ref, refmask, win, bands, x0=0):
    """pano_rows / mask_rows: canvas rows [rows_y0, rows_y0 + n) x columns [x0, x0 + width) of the library's output.
    Compares what lies at least margin(bands) inside the window. -> (pixels compared, pixels differing, max |diff|)."""
    wx, wy, ww, wh = win
    g = margin(bands)
    ya, yb = max(wy + g, rows_y0), min(wy + wh - g, rows_y0 + pano_rows.shape[0])
    xa, xb = max(wx + g, x0), min(wx + ww - g, x0 + pano_rows.shape[1])
    if ya >= yb or xa >= xb:
        return 0, 0, 0
    got = pano_rows[ya - rows_y0:yb - rows_y0, xa - x0:xb - x0]
    gm = mask_rows[ya - rows_y0:yb - rows_y0, xa - x0:xb - x0] if mask_rows is not None else None
    want = ref[ya - wy:yb - wy, xa - wx:xb - wx]
    wm = refmask[ya - wy:yb - wy, xa - wx:xb - wx]
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    bad = int(np.count_nonzero(d.max(axis=2)))
    if gm is not None:
        bad += int(np.count_nonzero(gm != wm))
    return int(got.shape[0] * got.shape[1]), bad, int(d.max()) if d.size else 0
