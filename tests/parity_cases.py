"""Parity cases shared by the CPU-emulation tests (logic check, no GPU) and the GPU tests (the
parity tests proper). Every case runs the library under test through the C ABI and compares with the
oracle on the same seeded inputs. Integer / index work must be bit-exact; the blended uint8 output is
held to the north-star bar (<= 1 LSB on >= 99.99 %) and additionally expected bit-exact."""
import os
import numpy as np

from drone_image_stitch_cpp_b200 import _lib as L
from drone_image_stitch_cpp_b200 import compositor as CP
from drone_image_stitch_cpp_b200 import synth
from oracle import ds_oracle as O

from helpers import assert_blend_parity


def oracle_warp(spec):
    """spec: dict(img, kind, ...) -> corner, warped, mask, xy, a (oracle side)."""
    img = spec["img"]
    h, w = img.shape[:2]
    if spec["kind"] == "plane":
        r = O.warp_frame(img, spec["K"], spec["R"], spec["scale"], spec.get("affine", True))
        corner, warped, mask, xy, a = r["corner"], r["warped"], r["mask"], r["xy"], r["a"]
    elif spec["kind"] == "affine":
        dw, dh = spec["size"]
        xy, a = O.affine_tables(spec["M"], dw, dh)
        warped = O.remap_bilinear(img, xy, a, "constant")
        mask = O.affine_nearest_mask(spec["M"], dw, dh, w, h)
        corner = spec["corner"]
    else:
        dw, dh = spec["size"]
        xy, a = O.persp_tables(spec["M"], dw, dh)
        warped = O.remap_bilinear(img, xy, a, "constant")
        mask = O.persp_nearest_mask(spec["M"], dw, dh, w, h)
        corner = spec["corner"]
    if spec.get("gain") is not None:          # applyChannelGainInPlace: float32
        g = np.asarray(spec["gain"], np.float32)
        warped = np.clip(np.rint(warped.astype(np.float32) * g[None, None, :]), 0, 255).astype(np.uint8)
    if spec.get("cgain") is not None:         # ExposureCompensator::apply, scalar gains: float64 (pinned vs cv2.multiply)
        g = np.asarray(spec["cgain"], np.float64)
        warped = np.clip(np.rint(warped.astype(np.float64) * g[None, None, :]), 0, 255).astype(np.uint8)
    if spec.get("gain_map") is not None:      # BlocksGainCompensator::apply: float32 per pixel
        warped = np.clip(np.rint(warped.astype(np.float32) * spec["gain_map"][:, :, None]), 0, 255).astype(np.uint8)
    if spec.get("gain_blocks") is not None:   # the same from the compensator's block map: resize(INTER_LINEAR) to the bbox first
        gm = O.resize_linear_f32(spec["gain_blocks"], warped.shape[1], warped.shape[0])
        warped = np.clip(np.rint(warped.astype(np.float32) * gm[:, :, None]), 0, 255).astype(np.uint8)
    if spec.get("global_stage"):
        # stitchInterStripsCustom (stitch_global.cpp:479-486, :643-658): content mask, seam mask resized with
        # INTER_NEAREST + threshold, soft blend mask; the blender is fed with the soft mask itself
        content = O.content_mask(img, spec["M"], dw, dh)
        seam = None
        if spec.get("seam_lowres") is not None:
            seam = O.threshold_gt1(O.resize_nearest(spec["seam_lowres"], dw, dh))
        mask = O.soft_blend_mask(seam, content, spec.get("sigma", 10.0))
    elif spec.get("seam_lowres") is not None:
        mask = mask & O.seam_mask_upsize(spec["seam_lowres"], mask.shape[1], mask.shape[0])
    elif spec.get("seam") is not None:
        mask = mask & spec["seam"]
    return corner, warped, mask, xy, a


def lib_transform(spec):
    if spec["kind"] == "plane":
        return CP.plane_transform(spec["K"], spec["R"], spec["scale"], spec.get("affine", True))
    if spec["kind"] == "affine":
        return CP.affine_transform(spec["M"], spec["corner"], spec["size"])
    return CP.homography_transform(spec["M"], spec["corner"], spec["size"])


def run_case(lib, specs, blend, bands, check_taps=True, out_format="bgr", band_split=None, exact=True):
    ow = [oracle_warp(s) for s in specs]
    corners = [o[0] for o in ow]
    sizes = [(o[1].shape[1], o[1].shape[0]) for o in ow]
    roi = O.result_roi(corners, sizes)
    bl = O.MultiBand(roi, bands) if blend == "multiband" else O.Feather(roi, 0.02)
    taps = []
    for o in ow:
        if blend == "multiband":
            taps.append(bl.feed(o[1].astype(np.int16), o[2], o[0], taps=check_taps))
        else:
            bl.feed(o[1].astype(np.int16), o[2], o[0])
    ref16, refmask = bl.blend()
    ref = O.s16_to_u8(ref16)

    xfs = [lib_transform(s) for s in specs]
    rois = [CP.warp_roi(xf, s["img"].shape[1], s["img"].shape[0], lib) for xf, s in zip(xfs, specs)]
    for r, c, sz in zip(rois, corners, sizes):
        assert (r[0], r[1]) == tuple(c) and (r[2], r[3]) == tuple(sz), (r, c, sz)
    assert CP.result_roi(rois) == roi

    def make(band=None, slice_rows=0):
        """slice_rows > 0: the pipelined schedule - asynchronous uploads, composite in row slices of that height,
        download slice by slice."""
        # small upload chunks for the pipelined schedule, so that every slice waits for its own source rows only
        old = os.environ.get("DS_UPLOAD_CHUNK_ROWS")
        if slice_rows > 0:
            os.environ["DS_UPLOAD_CHUNK_ROWS"] = "24"
        try:
            cv = CP.Canvas(roi, blend, bands, 0.02, out_format, 0, band=band, lib=lib, pipeline_rows=slice_rows)
        finally:
            if old is None:
                os.environ.pop("DS_UPLOAD_CHUNK_ROWS", None)
            else:
                os.environ["DS_UPLOAD_CHUNK_ROWS"] = old
        for i, (s, xf) in enumerate(zip(specs, xfs)):
            if band is not None and not cv.touches(rois[i]):
                continue
            gs = bool(s.get("global_stage"))
            cv.upload(i, s["img"], xf, seam_mask=s.get("seam"), channel_gain=s.get("gain"), seam_lowres=s.get("seam_lowres"),
                      compensator_gain=s.get("cgain"), gain_map=s.get("gain_map"), gain_blocks=s.get("gain_blocks"), async_=slice_rows > 0,
                      content_mask=gs, seam_nearest=gs and s.get("seam_lowres") is not None,
                      soft_mask=(s.get("sigma", 10.0) if gs else None))
        if slice_rows > 0:
            cv.composite_async()
        else:
            cv.composite()
        return cv

    cv = make()
    info = cv.info()
    if blend == "multiband":
        assert info.num_bands == bl.bands
        assert (info.padded_width, info.padded_height) == (bl.roi[2], bl.roi[3])
    pano, mask = cv.download()
    if out_format == "bgra":
        mask = pano[:, :, 3].copy()
        pano = np.ascontiguousarray(pano[:, :, :3])
    if check_taps:
        for i, o in enumerate(ow):
            xy, a = cv.maps(i)
            assert np.array_equal(xy, o[3]) and np.array_equal(a, o[4]), f"frame {i}: INTER_BITS tables differ"
            img, mk = cv.warped(i)
            assert np.array_equal(img, o[1]), f"frame {i}: warped image differs"
            assert np.array_equal(mk, o[2]), f"frame {i}: warped mask differs"
            if blend == "multiband":
                g, gs, ws = taps[i]
                for l in range(1, bl.bands + 1):
                    G, W, dims = cv.frame_level(i, l)
                    assert np.array_equal(G, gs[l]), f"frame {i} level {l}: Gaussian level differs"
                    assert np.array_equal(W, ws[l]), f"frame {i} level {l}: weight level differs"
    assert np.array_equal(mask, refmask), "result mask differs"
    st = assert_blend_parity(pano, ref, exact=exact)
    # the pipelined schedule (row slices chasing the uploads) must give the same bytes, per-frame pyramids included
    for slice_rows in sorted({max(1 << info.num_bands, roi[3] // 5), max(1 << info.num_bands, roi[3] // 2)}):
        cs = make(slice_rows=slice_rows)
        ps, ms = cs.download()
        cs.synchronize()
        if out_format == "bgra":
            ms = ps[:, :, 3].copy()
            ps = np.ascontiguousarray(ps[:, :, :3])
        assert np.array_equal(ps, pano) and np.array_equal(ms, mask), f"sliced composite ({slice_rows} rows) differs"
        if check_taps and blend == "multiband":
            for i in range(len(specs)):
                for l in range(1, bl.bands + 1):
                    G, W, _ = cs.frame_level(i, l)
                    G0, W0, _ = cv.frame_level(i, l)
                    assert np.array_equal(G, G0) and np.array_equal(W, W0), f"sliced: frame {i} level {l} differs"
        cs.close()
    if band_split:
        # virtual bands: K handles over row bands must reproduce the single-handle result exactly
        m = 1 << info.num_bands
        H = info.padded_height
        edges = sorted(set([0] + [min(H, ((H * k // band_split) // m) * m) for k in range(1, band_split)] + [H]))
        rows = []
        for y0, y1 in zip(edges[:-1], edges[1:]):
            if y0 >= roi[3]:
                continue
            cb = make(band=(y0, y1), slice_rows=(1 << info.num_bands) * 2 if len(rows) % 2 else 0)
            cb.synchronize()
            pb, mb_ = cb.download()
            if out_format == "bgra":
                pb = np.ascontiguousarray(pb[:, :, :3])
            rows.append(pb)
            cb.close()
        stacked = np.concatenate(rows, axis=0)
        assert np.array_equal(stacked, pano), "banded result differs from the single-band result"
    cv.close()
    return st


def windowed_oracle(frames, Ks, Rs, scale, bands, roi, win):
    """SURVEY 8(c) P17, the oracle for canvases too large to blend whole on the CPU: the blender is prepared on an aligned
    window of the canvas (origin a multiple of 2^bands from the canvas origin), every warped frame is cropped to the
    window, and pixels at least 8 * 2^bands inside the window equal the full-canvas result bit for bit.
    win = (x, y, w, h) relative to the canvas origin; -> (pano16 -> u8, mask) of the window."""
    wx, wy, ww, wh = win
    m = 1 << bands
    assert wx % m == 0 and wy % m == 0 and ww % m == 0 and wh % m == 0
    ax, ay = roi[0] + wx, roi[1] + wy
    bl = O.MultiBand((ax, ay, ww, wh), bands)
    assert bl.bands == bands
    for f, K, R in zip(frames, Ks, Rs):
        w = O.warp_frame(f, K, R, scale)
        cx, cy = w["corner"]
        bw, bh = w["size"]
        x0, y0, x1, y1 = max(cx, ax), max(cy, ay), min(cx + bw, ax + ww), min(cy + bh, ay + wh)
        if x0 >= x1 or y0 >= y1:
            continue
        img = np.ascontiguousarray(w["warped"][y0 - cy:y1 - cy, x0 - cx:x1 - cx]).astype(np.int16)
        msk = np.ascontiguousarray(w["mask"][y0 - cy:y1 - cy, x0 - cx:x1 - cx])
        bl.feed(img, msk, (x0, y0))
    ref16, refmask = bl.blend()
    return O.s16_to_u8(ref16), refmask


def check_window(pano, mask, ref, refmask, win, bands):
    wx, wy, ww, wh = win
    g = 8 << bands
    sl = (slice(wy + g, wy + wh - g), slice(wx + g, wx + ww - g))
    rs = (slice(g, wh - g), slice(g, ww - g))
    assert np.array_equal(mask[sl], refmask[rs]), "result mask differs inside the window"
    return assert_blend_parity(pano[sl], ref[rs])


def case_windowed_oracle(lib):
    # the windowed oracle itself, at a size where the full-canvas oracle exists: both must agree with the library
    sv = synth.grid_survey(3, 3, 400, 300, overlap=0.6, seed=17, rot_deg=2.0)
    bands = 3
    pano, mask, roi = CP.compose_panorama(sv.frames, sv.Ks, sv.Rs, sv.scale, "multiband", bands, lib=lib)
    full, fullmask, roi2 = O.compose_port(sv.frames, sv.Ks, sv.Rs, sv.scale, "multiband", bands)
    assert roi == roi2 and np.array_equal(pano, full) and np.array_equal(mask, fullmask)
    win = (128, 96, 384, 320)
    ref, refmask = windowed_oracle(sv.frames, sv.Ks, sv.Rs, sv.scale, bands, roi, win)
    check_window(pano, mask, ref, refmask, win, bands)
    # oracle/windowed.py (frames warped only over their part of the window; what bench.py's in-run check uses) is the same oracle
    from oracle import windowed as W
    ref2, refmask2 = W.compose_window(sv.frames, sv.Ks, sv.Rs, sv.scale, bands, roi, win)
    assert np.array_equal(ref2, ref) and np.array_equal(refmask2, refmask)
    n, bad, mx = W.compare_inside(pano[100:300], mask[100:300], 100, ref2, refmask2, win, bands)
    assert n > 0 and bad == 0 and mx == 0


def plane_specs(sv):
    return [dict(kind="plane", img=f, K=K, R=R, scale=sv.scale) for f, K, R in zip(sv.frames, sv.Ks, sv.Rs)]


def affine_specs(seed=11, n=3, fw=360, fh=260, homography=False):
    """Frames placed like stitchInterStripsCustom does (stitch_global.cpp:439-480): forward double
    transform, bbox by floor/ceil of the transformed corners, translation made bbox-relative."""
    rng = np.random.default_rng(seed)
    ortho = synth.orthophoto(fh * 2 + 200, fw * 2 + 300, seed).numpy()
    specs = []
    for i in range(n):
        th, s = rng.uniform(-0.15, 0.15), rng.uniform(0.9, 1.1)
        tx, ty = 40 + i * fw * 0.45 + rng.uniform(-10, 10), 30 + (i % 2) * fh * 0.4 + rng.uniform(-10, 10)
        Hm = np.array([[s * np.cos(th), -s * np.sin(th), tx], [s * np.sin(th), s * np.cos(th), ty], [0, 0, 1.0]])
        if homography:
            Hm[2, 0], Hm[2, 1] = rng.uniform(-5e-5, 5e-5), rng.uniform(-5e-5, 5e-5)
        pts = np.array([[0, 0, 1], [fw, 0, 1], [fw, fh, 1], [0, fh, 1]], np.float64).T
        q = Hm @ pts
        q = q[:2] / q[2:]
        x0, y0 = int(np.floor(q[0].min())), int(np.floor(q[1].min()))
        bw, bh = max(1, int(np.ceil(q[0].max())) - x0), max(1, int(np.ceil(q[1].max())) - y0)
        Tm = np.array([[1, 0, -x0], [0, 1, -y0], [0, 0, 1.0]])
        M = Tm @ Hm
        img = np.ascontiguousarray(ortho[10 * i:10 * i + fh, 20 * i:20 * i + fw])
        specs.append(dict(kind="homography" if homography else "affine", img=img, M=(M if homography else M[:2]),
                          corner=(x0, y0), size=(bw, bh)))
    return specs


# ---- the case list: name -> callable(lib)

def case_small_feather(lib):
    return run_case(lib, plane_specs(synth.grid_survey(3, 2, 400, 300, overlap=0.6, seed=1, work_scale=0.37)), "feather", 0, band_split=3)


def case_small_mb5(lib):
    return run_case(lib, plane_specs(synth.grid_survey(3, 2, 400, 300, overlap=0.6, seed=1, work_scale=0.37)), "multiband", 5, band_split=2)


def case_small_mb3_bgra(lib):
    return run_case(lib, plane_specs(synth.grid_survey(2, 2, 333, 251, overlap=0.5, seed=4)), "multiband", 3, out_format="bgra", band_split=3)


def case_mb1_and_mb0(lib):
    sv = synth.grid_survey(2, 1, 200, 150, overlap=0.5, seed=5)
    run_case(lib, plane_specs(sv), "multiband", 1)
    return run_case(lib, plane_specs(sv), "multiband", 0)


def case_mb8_crops_bands(lib):
    # 8 requested bands on a small canvas: exercises the gap clamp / shift-back of the feed ROI
    return run_case(lib, plane_specs(synth.grid_survey(2, 2, 300, 220, overlap=0.7, seed=6, work_scale=0.5)), "multiband", 8)


def case_single_frame(lib):
    sv = synth.grid_survey(1, 1, 257, 131, seed=7)
    run_case(lib, plane_specs(sv), "feather", 0)
    return run_case(lib, plane_specs(sv), "multiband", 5)


def case_big_rotation(lib):
    sv = synth.grid_survey(3, 1, 300, 200, overlap=0.5, seed=8, rot_deg=25.0, scale_jit=0.1)
    run_case(lib, plane_specs(sv), "feather", 0)
    return run_case(lib, plane_specs(sv), "multiband", 4)


def case_plane_warper_general_K(lib):
    rng = np.random.default_rng(9)
    sv = synth.grid_survey(2, 1, 240, 180, overlap=0.5, seed=9)
    specs = plane_specs(sv)
    for s in specs:
        s["affine"] = False
        s["K"] = np.array([[1.0, 0, 3.5], [0, 1.02, -2.25], [0, 0, 1]], np.float32)
        th = rng.uniform(-0.02, 0.02)
        s["R"] = np.array([[np.cos(th), -np.sin(th), 0.01], [np.sin(th), np.cos(th), -0.02], [1e-5, -2e-5, 1.0]], np.float32)
        s["scale"] = 1.0
    return run_case(lib, specs, "multiband", 3)


def case_affine_f64(lib):
    run_case(lib, affine_specs(11), "multiband", 5)
    return run_case(lib, affine_specs(12), "feather", 0)


def case_homography_f64(lib):
    return run_case(lib, affine_specs(13, homography=True), "multiband", 4)


def case_seam_and_gain(lib):
    sv = synth.grid_survey(2, 2, 260, 200, overlap=0.6, seed=14)
    specs = plane_specs(sv)
    rng = np.random.default_rng(14)
    for i, s in enumerate(specs):
        o = O.warp_frame(s["img"], s["K"], s["R"], s["scale"])
        bw, bh = o["size"]
        seam = np.full((bh, bw), 255, np.uint8)
        seam[:, : bw // 3] = rng.integers(0, 256, (bh, bw // 3)).astype(np.uint8) if i % 2 else 0
        s["seam"] = seam
        s["gain"] = (1.0 + 0.1 * i, 0.95, 1.2 - 0.1 * i)
    run_case(lib, specs, "multiband", 4)
    return run_case(lib, specs, "feather", 0)


def case_medium_mb3_interior(lib):
    # frames large against the tile size: exercises the border-free / uniform-mask fast paths of the level-0 kernel
    return run_case(lib, plane_specs(synth.grid_survey(2, 2, 900, 700, overlap=0.6, seed=21, work_scale=0.45)), "multiband", 3, band_split=2)


def case_many_frames_one_spot(lib):
    # 70 frames over the same spot: more than the 64 frames per tile the packed-lane kernels accept, so the
    # library must route every level through the generic kernels (int16 wrap-around semantics preserved)
    rng = np.random.default_rng(70)
    base = synth.orthophoto(160, 200, 70).numpy()
    specs = []
    for i in range(70):
        th = rng.uniform(-0.05, 0.05)
        R = np.array([[np.cos(th), -np.sin(th), rng.uniform(-6, 6)], [np.sin(th), np.cos(th), rng.uniform(-6, 6)], [0, 0, 1]], np.float32)
        img = np.ascontiguousarray(base[rng.integers(0, 20):, rng.integers(0, 20):][:120, :150])
        specs.append(dict(kind="plane", img=img, K=np.eye(3, dtype=np.float32), R=R, scale=1.0))
    return run_case(lib, specs, "multiband", 2, check_taps=False)


def case_serpentine_strip_scaled(lib):
    # BASELINE config 3 layout at 1/6 linear scale: 3 flight lines x 12 frames, 70 % forward / 32 % side overlap
    sv = synth.grid_survey(12, 3, 912, 608, overlap=0.7, side_overlap=0.32, seed=303)
    return run_case(lib, plane_specs(sv), "multiband", 5, check_taps=False, band_split=4)


def case_seam_lowres_upsizing(lib):
    # the per-frame seam-mask hand-off of composePanorama: low-res seam mask (seam_estimation_resol) ->
    # dilate -> LINEAR_EXACT resize -> AND with the warped mask
    sv = synth.grid_survey(2, 2, 300, 220, overlap=0.6, seed=88)
    specs = plane_specs(sv)
    rng = np.random.default_rng(88)
    for i, s in enumerate(specs):
        lw, lh = 37 + 3 * i, 29 + 2 * i
        m = np.zeros((lh, lw), np.uint8)
        m[rng.integers(0, 5):lh - rng.integers(0, 5), rng.integers(0, 6):lw - rng.integers(0, 6)] = 255
        m[lh // 2:lh // 2 + 3, : lw // 3] = 0
        s["seam_lowres"] = m
    run_case(lib, specs, "multiband", 4)
    return run_case(lib, specs, "feather", 0)


def case_exposure_gains(lib):
    # the three gain applications of the reference between warp and blend: per-strip channel gain (float32),
    # compensator scalar gains (float64) and the BlocksGain per-pixel map (float32)
    sv = synth.grid_survey(2, 2, 280, 210, overlap=0.6, seed=91, work_scale=0.5)
    specs = plane_specs(sv)
    rng = np.random.default_rng(91)
    for i, s in enumerate(specs):
        o = O.warp_frame(s["img"], s["K"], s["R"], s["scale"])
        bw, bh = o["size"]
        if i != 1:
            s["gain_map"] = (rng.random((bh, bw)) * 0.5 + 0.75).astype(np.float32)
        if i != 2:
            s["cgain"] = (1.003 + 0.01 * i, 0.997, 1.1)
        if i == 0:
            s["gain"] = (1.07, 0.93, 1.21)
    run_case(lib, specs, "multiband", 3)
    run_case(lib, specs, "feather", 0)
    # BlocksGainCompensator as the reference configures it (stitch_robust.cpp:209-211): the compensator's 32x32-block gain
    # maps go in as they are, the f32 INTER_LINEAR resize to the warped size happens on the device
    for i, s in enumerate(specs):
        s.pop("gain_map", None)
        o = O.warp_frame(s["img"], s["K"], s["R"], s["scale"])
        bw, bh = o["size"]
        s["gain_blocks"] = (rng.random(((bh + 31) // 32, (bw + 31) // 32)) * 0.6 + 0.7).astype(np.float32)
    return run_case(lib, specs, "multiband", 3)


def global_stage_specs(seed=31, n=3, fw=420, fh=300):
    """Strip panoramas as the global stage sees them: black wedges / holes left by the strip stage, placed by
    transformedBoundingRect (affine_specs), with a low-resolution seam mask per strip."""
    specs = affine_specs(seed, n, fw, fh)
    rng = np.random.default_rng(seed)
    for i, s in enumerate(specs):
        img = s["img"].copy()
        yy, xx = np.mgrid[0:fh, 0:fw]
        img[(yy < 0.12 * xx - 8 * i) | (yy > fh - 14 + 0.04 * xx)] = 0            # wedges
        img[fh // 3:fh // 3 + 11, fw // 2:fw // 2 + 37] = rng.integers(0, 5, (11, 37, 3), dtype=np.uint8)   # near-black hole
        s["img"] = img
        bw, bh = s["size"]
        lw, lh = max(4, bw // 7), max(4, bh // 7)
        m = np.full((lh, lw), 255, np.uint8)
        yy, xx = np.mgrid[0:lh, 0:lw]
        if i % 2:
            m[xx > 0.6 * lw + 0.2 * yy] = 0
        else:
            m[xx < 0.35 * lw - 0.1 * yy] = 0
        m[rng.integers(0, lh), rng.integers(0, lw)] = 1     # a value ensureBinaryMask drops
        s["seam_lowres"] = m
        s["global_stage"] = True
        if i:
            s["gain"] = (1.0 + 0.04 * i, 0.97, 1.05)
            s["cgain"] = (0.98, 1.02 + 0.01 * i, 1.0)
    return specs


def case_global_stage_masks(lib):
    # buildWarpedContentMask + NEAREST seam resize + buildSoftBlendMask on the device, fed to the
    # multi-band blender like stitchInterStripsCustom does; also the two-step flow of the reference (warp first, masks and
    # gains once the CPU-side exposure / seam steps are done) through ds_update_frame_opts
    specs = global_stage_specs()
    run_case(lib, specs, "multiband", 5, check_taps=False)
    ow = [oracle_warp(s) for s in specs]
    roi = O.result_roi([o[0] for o in ow], [(o[1].shape[1], o[1].shape[0]) for o in ow])
    cv = CP.Canvas(roi, "multiband", 5, lib=lib)
    for i, s in enumerate(specs):
        cv.upload(i, s["img"], lib_transform(s), content_mask=True)
        dw, dh = s["size"]
        assert np.array_equal(cv.frame_mask(i, 1), O.content_mask(s["img"], s["M"], dw, dh)), f"strip {i}: content mask differs"
        img, _ = cv.warped(i)
        assert np.array_equal(img, O.remap_bilinear(s["img"], *O.affine_tables(s["M"], dw, dh), "constant"))
    for i, s in enumerate(specs):
        cv.update_opts(i, seam_lowres=s["seam_lowres"], seam_nearest=True, content_mask=True, soft_mask=True,
                       channel_gain=s.get("gain"), compensator_gain=s.get("cgain"))
        assert np.array_equal(cv.frame_mask(i, 0), ow[i][2]), f"strip {i}: soft blend mask differs"
        assert ((ow[i][2] > 0) & (ow[i][2] < 255)).any()
    cv.composite()
    pano, mask = cv.download()
    bl = O.MultiBand(roi, 5)
    for o in ow:
        bl.feed(o[1].astype(np.int16), o[2], o[0])
    ref16, refmask = bl.blend()
    assert np.array_equal(mask, refmask)
    assert_blend_parity(pano, O.s16_to_u8(ref16))
    # dropping the options again gives the plain composite
    for i in range(len(specs)):
        cv.update_opts(i)
    cv.composite()
    p2, _ = cv.download()
    plain = [dict(s, global_stage=False, seam_lowres=None, gain=None, cgain=None) for s in specs]
    cv2_ = CP.Canvas(roi, "multiband", 5, lib=lib)
    for i, s in enumerate(plain):
        cv2_.upload(i, s["img"], lib_transform(s))
    cv2_.composite()
    assert np.array_equal(p2, cv2_.download()[0])
    cv.close(); cv2_.close()


def case_seam_phase_warps(lib):
    # the seam-phase warps of composePanorama - seam-scale images (seam_est_resol), cameras scaled by
    # seam_work_aspect, LINEAR / REFLECT image warp and NEAREST / CONSTANT mask warp - and a strip's warpAffine, each as
    # one stand-alone call
    import cv2  # noqa: F401  (only to resize like the registration stage does; cv2 is a test dependency)
    sv = synth.grid_survey(2, 2, 640, 480, overlap=0.6, seed=55, work_scale=0.6)
    seam_work_aspect = 0.31
    for f, K, R in zip(sv.frames, sv.Ks, sv.Rs):
        small = cv2.resize(f, None, fx=seam_work_aspect, fy=seam_work_aspect, interpolation=cv2.INTER_LINEAR_EXACT)
        Ks = np.asarray(K, np.float32).copy()
        Ks[0, 0] *= seam_work_aspect; Ks[1, 1] *= seam_work_aspect; Ks[0, 2] *= seam_work_aspect; Ks[1, 2] *= seam_work_aspect
        scale = np.float32(sv.scale * seam_work_aspect)
        o = O.warp_frame(small, Ks, R, scale)
        corner, img, mask = CP.warp_frame(small, CP.plane_transform(Ks, R, scale), lib=lib)
        assert corner == tuple(o["corner"])
        assert np.array_equal(img, o["warped"]) and np.array_equal(mask, o["mask"])
    for s_ in affine_specs(56):
        c_, w_, m_, _, _ = oracle_warp(s_)
        corner, img, mask = CP.warp_frame(s_["img"], lib_transform(s_), lib=lib)
        assert corner == tuple(c_) and np.array_equal(img, w_) and np.array_equal(mask, m_)


def case_global_stage_edge_sizes(lib):
    # the mask chain and the crop on degenerate and awkward plane sizes: 1-pixel-wide / -high bboxes, widths around the
    # 8-column boundary of OpenCV's vector column filter, sizes around the 64-pixel tile of the blur kernel
    rng = np.random.default_rng(1)
    img = synth.orthophoto(120, 160, 9).numpy().copy()
    img[:20, :50] = 0
    M = np.array([[1.0, 0.03, -3.5], [-0.02, 1.0, -2.25]])
    for dw, dh in [(1, 1), (1, 40), (40, 1), (2, 2), (7, 200), (8, 200), (9, 200), (200, 3), (65, 65), (64, 64), (63, 129)]:
        low = (rng.random((max(1, dh // 3), max(1, dw // 3))) > 0.3).astype(np.uint8) * 255
        cv = CP.Canvas((0, 0, dw, dh), "multiband", 2, lib=lib)
        cv.upload(0, img, CP.affine_transform(M, (0, 0), (dw, dh)), content_mask=True, seam_lowres=low, seam_nearest=True, soft_mask=10.0)
        content = O.content_mask(img, M, dw, dh)
        soft = O.soft_blend_mask(O.threshold_gt1(O.resize_nearest(low, dw, dh)), content, 10.0)
        assert np.array_equal(cv.frame_mask(0, 1), content), (dw, dh)
        assert np.array_equal(cv.frame_mask(0, 0), soft), (dw, dh)
        cv.composite()
        pano, mask = cv.download()
        bl = O.MultiBand((0, 0, dw, dh), 2)
        bl.feed(O.remap_bilinear(img, *O.affine_tables(M, dw, dh), "constant").astype(np.int16), soft, (0, 0))
        r16, rm = bl.blend()
        assert np.array_equal(pano, O.s16_to_u8(r16)) and np.array_equal(mask, rm), (dw, dh)
        try:
            assert cv.auto_crop_rect() == O.auto_crop_rect(pano), (dw, dh)
        except L.DroneStitchError as e:
            assert e.code == L.DS_ERR_UNSUPPORTED
        cv.close()


def case_auto_crop(lib):
    # autoCropBlackBorder's rectangle computed from the canvas in device memory
    rng = np.random.default_rng(41)
    sv = synth.grid_survey(2, 2, 300, 220, overlap=0.5, seed=41, rot_deg=12.0)
    specs = plane_specs(sv)
    # a small far-away frame: a second external contour that must lose; dark pixels inside the content
    small = np.ascontiguousarray(specs[0]["img"][:40, :50])
    R = np.array([[1, 0, 900.0], [0, 1, -60.0], [0, 0, 1]], np.float32)
    specs.append(dict(kind="plane", img=small, K=np.eye(3, dtype=np.float32), R=R, scale=sv.scale))
    for s_ in specs[:4]:
        img = s_["img"].copy()
        img[rng.integers(0, img.shape[0], 200), rng.integers(0, img.shape[1], 200)] = 0
        s_["img"] = img
    for blend, bands in (("multiband", 3), ("feather", 0)):
        xfs = [lib_transform(s_) for s_ in specs]
        rois = [CP.warp_roi(xf, s_["img"].shape[1], s_["img"].shape[0], lib) for xf, s_ in zip(xfs, specs)]
        roi = CP.result_roi(rois)
        cv = CP.Canvas(roi, blend, bands, lib=lib)
        for i, (s_, xf) in enumerate(zip(specs, xfs)):
            cv.upload(i, s_["img"], xf)
        cv.composite()
        pano, _ = cv.download()
        rect = cv.auto_crop_rect()
        assert rect == O.auto_crop_rect(pano), (rect, O.auto_crop_rect(pano))
        assert rect[2] < roi[2]   # the far-away frame is cropped off
        x, y, w, h = rect
        assert np.array_equal(cv.download(x, y, w, h)[0], pano[y:y + h, x:x + w])   # pano(max_rect).clone()
        cv.close()
    # BGRA canvas, a single frame
    cv = CP.Canvas(rois[0], "multiband", 2, out_format="bgra", lib=lib)
    cv.upload(0, specs[0]["img"], xfs[0])
    cv.composite()
    assert cv.auto_crop_rect() == O.auto_crop_rect(cv.download()[0])
    cv.close()
    # two regions of the same size: the bounds cannot tell which contour is larger -> refused, not guessed
    twins = [dict(specs[0]), dict(specs[0])]
    twins[1]["R"] = specs[0]["R"].copy(); twins[1]["R"][0, 2] += 700
    xfs = [lib_transform(s_) for s_ in twins]
    rois = [CP.warp_roi(xf, 300, 220, lib) for xf in xfs]
    cv = CP.Canvas(CP.result_roi(rois), "feather", 0, lib=lib)
    for i in range(2):
        cv.upload(i, twins[i]["img"], xfs[i])
    cv.composite()
    try:
        cv.auto_crop_rect()
        raise AssertionError("ambiguous crop was not refused")
    except L.DroneStitchError as e:
        assert e.code == L.DS_ERR_UNSUPPORTED
    cv.close()
    # row-band handles are refused
    cb = CP.Canvas(CP.result_roi(rois), "multiband", 2, band=(0, 64), lib=lib)
    cb.upload(0, twins[0]["img"], xfs[0])
    cb.composite()
    try:
        cb.auto_crop_rect()
        raise AssertionError("band handle accepted")
    except L.DroneStitchError as e:
        assert e.code == L.DS_ERR_UNSUPPORTED
    cb.close()


def case_bands8_and_row_bands(lib):
    # 8 bands (BASELINE config 5's blend depth) on a canvas just large enough, split into 3 row bands
    sv = synth.grid_survey(2, 3, 520, 400, overlap=0.55, seed=95)
    return run_case(lib, plane_specs(sv), "multiband", 8, check_taps=False, band_split=3)


def case_very_wide_canvas(lib):
    # one flight line whose canvas is wider than 65 535 px (BASELINE configs 3-5 are 69 k - 135 k px wide):
    # 64-bit row offsets, large TMA coordinates, tile indices beyond 16 bits
    n, fw, fh = 240, 512, 96
    plan = synth.plan_grid(n, 1, fw, fh, overlap=0.4, seed=97, rot_deg=1.0, trans_jit=6.0)
    rng = np.random.default_rng(97)
    base = synth.orthophoto(fh, fw, 97).numpy()
    frames = [np.ascontiguousarray(np.roll(base, int(rng.integers(0, fw)), axis=1)) for _ in range(n)]
    specs = [dict(kind="plane", img=f, K=K, R=R, scale=plan.scale) for f, K, R in zip(frames, plan.Ks, plan.Rs)]
    rois = [CP.warp_roi(lib_transform(s), fw, fh, lib) for s in specs]
    assert CP.result_roi(rois)[2] > 65535
    return run_case(lib, specs, "multiband", 5, check_taps=False)


CASES = {
    "global_stage_edge_sizes": case_global_stage_edge_sizes,
    "windowed_oracle": case_windowed_oracle,
    "seam_phase_warps": case_seam_phase_warps,
    "auto_crop": case_auto_crop,
    "global_stage_masks": case_global_stage_masks,
    "exposure_gains": case_exposure_gains,
    "bands8_and_row_bands": case_bands8_and_row_bands,
    "very_wide_canvas": case_very_wide_canvas,
    "seam_lowres_upsizing": case_seam_lowres_upsizing,
    "medium_mb3_interior": case_medium_mb3_interior,
    "many_frames_one_spot": case_many_frames_one_spot,
    "small_feather": case_small_feather,
    "small_mb5": case_small_mb5,
    "small_mb3_bgra": case_small_mb3_bgra,
    "mb1_and_mb0": case_mb1_and_mb0,
    "mb8_crops_bands": case_mb8_crops_bands,
    "single_frame": case_single_frame,
    "big_rotation": case_big_rotation,
    "plane_warper_general_K": case_plane_warper_general_K,
    "affine_f64": case_affine_f64,
    "homography_f64": case_homography_f64,
    "seam_and_gain": case_seam_and_gain,
}
