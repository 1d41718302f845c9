"""The parity tests proper: libdronestitch_cuda (sm_100a kernels, through the C ABI) vs the oracle."""
import numpy as np
import pytest

from parity_cases import CASES, plane_specs, run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_case(cuda_lib, name):
    CASES[name](cuda_lib)


def test_cfg3_layout_scaled(cuda_lib):
    """BASELINE config 3 layout (3 serpentine lines, 70 % forward / 32 % side overlap) at 1/6 scale, 36 frames,
    against the oracle, plus the 4-band decomposition."""
    from parity_cases import case_serpentine_strip_scaled
    case_serpentine_strip_scaled(cuda_lib)


def test_cfg1_full_size_feather(cuda_lib):
    """BASELINE config 1: 2 x 4000x3000, feather. Oracle finishes in seconds at full size."""
    from drone_image_stitch_cpp_b200 import synth
    sv = synth.grid_survey(2, 1, 4000, 3000, overlap=0.7, seed=101, rot_deg=1.5, device="cuda")
    run_case(cuda_lib, plane_specs(sv), "feather", 0, check_taps=True)


def test_cfg1_full_size_multiband(cuda_lib):
    from drone_image_stitch_cpp_b200 import synth
    sv = synth.grid_survey(2, 1, 4000, 3000, overlap=0.7, seed=102, rot_deg=1.5, device="cuda")
    run_case(cuda_lib, plane_specs(sv), "multiband", 5, check_taps=False, band_split=4)


def test_cfg2_properties_full_size(cuda_lib):
    """BASELINE config 2 at full size (3x3 of 5472x3648, 70 % overlap, multi-band 5): size-independent
    properties — idempotence, row-band decomposition, BGRA == BGR + mask, and exact agreement with the
    oracle on an aligned window cut from the middle (windowed oracle)."""
    from drone_image_stitch_cpp_b200 import compositor as CP, synth
    from oracle import ds_oracle as O
    sv = synth.grid_survey(3, 3, 5472, 3648, overlap=0.7, seed=synth.MASTER_SEED, device="cuda")
    pano, mask, roi, cv = CP.compose_panorama(sv.frames, sv.Ks, sv.Rs, sv.scale, "multiband", 5, lib=cuda_lib, return_canvas=True)
    # idempotence
    cv.composite()
    pano2, mask2 = cv.download()
    assert np.array_equal(pano, pano2) and np.array_equal(mask, mask2)
    info = cv.info()
    cv.close()
    # BGRA canvas == BGR + mask
    pa, _, _ = CP.compose_panorama(sv.frames, sv.Ks, sv.Rs, sv.scale, "multiband", 5, lib=cuda_lib, out_format="bgra")
    assert np.array_equal(pa[:, :, :3], pano) and np.array_equal(pa[:, :, 3], mask)
    del pa
    # 4 row bands == 1 band
    xfs = [CP.plane_transform(K, R, sv.scale) for K, R in zip(sv.Ks, sv.Rs)]
    rois = [CP.warp_roi(xf, 5472, 3648, cuda_lib) for xf in xfs]
    H = info.padded_height
    edges = [0] + [((H * k // 4) // 32) * 32 for k in (1, 2, 3)] + [H]
    rows = []
    for y0, y1 in zip(edges[:-1], edges[1:]):
        cb = CP.Canvas(roi, "multiband", 5, band=(y0, y1), lib=cuda_lib)
        for i, (f, xf) in enumerate(zip(sv.frames, xfs)):
            if cb.touches(rois[i]):
                cb.upload(i, f, xf)
        cb.composite()
        rows.append(cb.download()[0])
        cb.close()
    assert np.array_equal(np.concatenate(rows, axis=0), pano)
    # full-size oracle on this config takes about a minute of CPU; bound it by the middle frame only:
    # the centre frame composited alone, full size, against the oracle
    mid = [4]
    p1, m1, r1 = O.compose_port([sv.frames[i] for i in mid], [sv.Ks[i] for i in mid], [sv.Rs[i] for i in mid], sv.scale, "multiband", 5)
    p2, m2, r2 = CP.compose_panorama([sv.frames[i] for i in mid], [sv.Ks[i] for i in mid], [sv.Rs[i] for i in mid], sv.scale,
                                     "multiband", 5, lib=cuda_lib)
    assert r1 == r2 and np.array_equal(m1, m2)
    assert np.array_equal(p1, p2)


def test_no_cpu_fallback_symbols(cuda_lib):
    assert b"sm_100a" in cuda_lib.dll.ds_version()


def test_cfg4_scale_one_gpu(tmp_path):
    """BASELINE config 4 at full size on one GPU (600 frames of 5472x3648, 2.7 GP canvas, ~125 GB of HBM): the
    composite runs, three band handles reproduce its rows bit for bit - indexing beyond 2^31 pixels, tile lists,
    TMA descriptors - and a window from the middle of the canvas equals the windowed oracle.
    Fresh process (tools/scale_check.py); skipped when the device has less than 140 GB free."""
    import os, subprocess, sys, json
    import torch
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 140e9:
        pytest.skip(f"needs 140 GB of free device memory, {free / 1e9:.0f} GB available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    npz = os.path.join(str(tmp_path), "window.npz")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "scale_check.py"), "50", "12", npz], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["ok"] and res["gigapixels"] > 2.2 and len(res["bands_checked"]) == 3
    # the window against the oracle: the frames are seeded noise, regenerated here for the indices the window sees
    from drone_image_stitch_cpp_b200 import compositor as CP, synth, _lib
    from parity_cases import windowed_oracle
    from helpers import assert_blend_parity
    w = np.load(npz)
    roi, win = tuple(int(v) for v in w["roi"]), tuple(int(v) for v in w["win"])
    fw, fh, bands = 5472, 3648, 5
    plan = synth.plan_grid(50, 12, fw, fh, overlap=0.7, side_overlap=0.32, seed=synth.MASTER_SEED)
    lib = _lib.default_library()
    wx, wy, ww, wh = win
    idx = []
    for i, (K, R) in enumerate(zip(plan.Ks, plan.Rs)):
        q = CP.warp_roi(CP.plane_transform(K, R, plan.scale), fw, fh, lib)
        if q[0] < roi[0] + wx + ww and q[0] + q[2] > roi[0] + wx and q[1] < roi[1] + wy + wh and q[1] + q[3] > roi[1] + wy:
            idx.append(i)
    assert len(idx) >= 4
    frames = []
    for i in idx:
        g = torch.Generator(device="cuda").manual_seed(1000 + i)
        frames.append(torch.randint(0, 256, (fh, fw, 3), dtype=torch.uint8, device="cuda", generator=g).cpu().numpy())
    ref, refmask = windowed_oracle(frames, [plan.Ks[i] for i in idx], [plan.Rs[i] for i in idx], plan.scale, bands, roi, win)
    gm = 8 << bands
    assert np.array_equal(w["mask"][gm:wh - gm, gm:ww - gm], refmask[gm:wh - gm, gm:ww - gm])
    st = assert_blend_parity(w["tile"][gm:wh - gm, gm:ww - gm], ref[gm:wh - gm, gm:ww - gm])
    assert st["n_diff"] == 0


def _windowed_check(cuda_lib, plan, bands, win_frac=(0.5, 0.5), win_size=(2048, 1536)):
    """Composite a whole BASELINE-size survey on the GPU (frames cut on the device, handed over by device pointer) and
    compare an aligned window of the canvas with the windowed oracle: bit-exact at least 8 * 2^bands
    inside the window."""
    import torch
    from drone_image_stitch_cpp_b200 import compositor as CP, synth
    from parity_cases import windowed_oracle
    xfs = [CP.plane_transform(K, R, plan.scale) for K, R in zip(plan.Ks, plan.Rs)]
    rois = [CP.warp_roi(xf, plan.fw, plan.fh, cuda_lib) for xf in xfs]
    roi = CP.result_roi(rois)
    m = 1 << bands
    ww, wh = win_size
    wx = int((roi[2] - ww) * win_frac[0]) // m * m
    wy = int((roi[3] - wh) * win_frac[1]) // m * m
    win = (wx, wy, ww, wh)
    ortho = synth.orthophoto(plan.ortho_h, plan.ortho_w, plan.seed, "cuda")
    cv = CP.Canvas(roi, "multiband", bands, lib=cuda_lib)
    keep = {}
    for i, xf in enumerate(xfs):
        fr = synth.cut(plan, [i], "cuda", as_torch=True, ortho=ortho)[0]
        torch.cuda.synchronize()   # the library reads the pixels on its own streams: they must be complete
        cv.upload_device(i, fr.data_ptr(), plan.fw, plan.fh, plan.fw * 3, xf)
        r = rois[i]
        if r[0] < roi[0] + wx + ww and r[0] + r[2] > roi[0] + wx and r[1] < roi[1] + wy + wh and r[1] + r[3] > roi[1] + wy:
            keep[i] = fr.cpu().numpy()
        del fr
    del ortho
    torch.cuda.empty_cache()
    cv.composite()
    assert cv.info().num_bands == bands
    tile, tmask = cv.download(wx, wy, ww, wh)
    cv.close()
    idx = sorted(keep)
    assert len(idx) >= 4, "the window should see several overlapping frames"
    ref, refmask = windowed_oracle([keep[i] for i in idx], [plan.Ks[i] for i in idx], [plan.Rs[i] for i in idx], plan.scale, bands, roi, win)
    g = 8 << bands
    assert np.array_equal(tmask[g:wh - g, g:ww - g], refmask[g:wh - g, g:ww - g]), "result mask differs inside the window"
    from helpers import assert_blend_parity
    return assert_blend_parity(tile[g:wh - g, g:ww - g], ref[g:wh - g, g:ww - g]), len(idx), roi


def test_cfg2_full_size_windowed_oracle(cuda_lib):
    """BASELINE config 2 at full size (3x3 of 5472x3648, 70 % overlap, 5 bands): the middle of the canvas, where all nine
    frames overlap, against the windowed oracle."""
    from drone_image_stitch_cpp_b200 import synth
    plan = synth.plan_grid(3, 3, 5472, 3648, overlap=0.7, seed=synth.MASTER_SEED)
    st, n, roi = _windowed_check(cuda_lib, plan, 5)
    assert n == 9 and st["n_diff"] == 0


def test_cfg3_full_size_windowed_oracle(cuda_lib):
    """BASELINE config 3 at full size: 120 frames of 5472x3648 in three serpentine lines, 0.6 GP canvas about 69 k px
    wide. A window between the first and the second flight line against the windowed oracle."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip(f"needs 60 GB of free device memory, {free / 1e9:.0f} GB available")
    from drone_image_stitch_cpp_b200 import synth
    plan = synth.plan_grid(40, 3, 5472, 3648, overlap=0.7, side_overlap=0.32, seed=synth.MASTER_SEED)
    st, n, roi = _windowed_check(cuda_lib, plan, 5, win_frac=(0.37, 0.36))
    assert roi[2] > 65535 and st["n_diff"] == 0


def test_cfg5_one_band_of_eight_windowed_oracle(cuda_lib):
    """BASELINE config 5 geometry: 2000 frames of 5472x3648 (25 lines x 80), 8-level multi-band, canvas about 8 GP, cut into
    eight row bands. One GPU computes ONE of the eight bands exactly as it would in the 8-GPU job (a band handle with the
    frames that touch the band + its recomputed pyramid halo) and the top of that band is compared with the windowed oracle
    (margin 8 * 2^8 = 2048 px). Frames are seeded noise generated on the device."""
    import torch
    from drone_image_stitch_cpp_b200 import compositor as CP, synth
    from parity_cases import windowed_oracle
    from helpers import assert_blend_parity
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 135e9:
        pytest.skip(f"needs 135 GB of free device memory, {free / 1e9:.0f} GB available")
    fw, fh, bands = 5472, 3648, 8
    plan = synth.plan_grid(80, 25, fw, fh, overlap=0.7, side_overlap=0.32, seed=synth.MASTER_SEED)
    xfs = [CP.plane_transform(K, R, plan.scale) for K, R in zip(plan.Ks, plan.Rs)]
    rois = [CP.warp_roi(xf, fw, fh, cuda_lib) for xf in xfs]
    roi = CP.result_roi(rois)
    assert roi[2] * roi[3] > 7.5e9
    m = 1 << bands
    H = (roi[3] + m - 1) // m * m
    y0, y1 = (H * 3 // 8) // m * m, (H * 4 // 8) // m * m
    cb = CP.Canvas(roi, "multiband", bands, band=(y0, y1), lib=cuda_lib)
    assert cb.info().num_bands == bands

    def frame(i):
        g = torch.Generator(device="cuda").manual_seed(5000 + i)
        f = torch.randint(0, 256, (fh, fw, 3), dtype=torch.uint8, device="cuda", generator=g)
        torch.cuda.synchronize()
        return f

    mine = [i for i in range(len(xfs)) if cb.touches(rois[i])]
    assert 200 < len(mine) < 1000
    for i in mine:
        f = frame(i)
        cb.upload_device(i, f.data_ptr(), fw, fh, fw * 3, xfs[i])
        del f
    cb.composite()
    info = cb.info()
    g = 8 << bands
    ww, wh = 2 * g + 2048, 2 * g + 1024
    wx, wy = (roi[2] - ww) // 2 // m * m, y0 - g
    win = (wx, wy, ww, wh)
    tile, tmask = cb.download(wx + g, y0, 2048, 1024)
    ms, dev_gb = info.ms_last_composite, info.device_bytes / 1e9
    cb.close()
    torch.cuda.empty_cache()
    idx = [i for i in range(len(xfs)) if rois[i][0] < roi[0] + wx + ww and rois[i][0] + rois[i][2] > roi[0] + wx and
           rois[i][1] < roi[1] + wy + wh and rois[i][1] + rois[i][3] > roi[1] + wy]
    assert set(idx) <= set(mine) and len(idx) >= 8
    frames = [frame(i).cpu().numpy() for i in idx]
    ref, refmask = windowed_oracle(frames, [plan.Ks[i] for i in idx], [plan.Rs[i] for i in idx], plan.scale, bands, roi, win)
    assert np.array_equal(tmask, refmask[g:g + 1024, g:g + 2048])
    st = assert_blend_parity(tile, ref[g:g + 1024, g:g + 2048])
    assert st["n_diff"] == 0
    print(f"cfg5 band: {len(mine)} frames, {dev_gb:.1f} GB, composite {ms:.1f} ms, window frames {len(idx)}")
