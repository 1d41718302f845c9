"""Randomised parity cases: random survey geometry, transform kind, blend, bands, row-band split."""
import numpy as np

import parity_cases as P
from drone_image_stitch_cpp_b200 import synth


def random_case(lib, seed):
    rng = np.random.default_rng(seed)
    nx, ny = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    fw, fh = int(rng.integers(24, 420)), int(rng.integers(24, 320))
    ov = float(rng.uniform(0.2, 0.85))
    rot = float(rng.choice([0.0, 1.0, 3.0, 12.0, 40.0]))
    sj = float(rng.choice([0.0, 0.02, 0.2]))
    ws = float(rng.choice([1.0, 0.45, 0.37, 2.0]))
    blend = "multiband" if rng.random() < 0.75 else "feather"
    bands = int(rng.integers(0, 8))
    split = int(rng.choice([0, 2, 3]))
    kind = str(rng.choice(["plane", "plane", "plane", "affine", "homography"]))
    desc = f"seed={seed} {kind} {nx}x{ny} {fw}x{fh} ov={ov:.2f} rot={rot} sj={sj} ws={ws} {blend} bands={bands} split={split}"
    if kind == "plane":
        sv = synth.grid_survey(nx, ny, fw, fh, overlap=ov, seed=seed, rot_deg=rot, scale_jit=sj, work_scale=ws,
                               trans_jit=float(rng.uniform(0, 25)))
        specs = P.plane_specs(sv)
        if rng.random() < 0.3:
            for s in specs:
                s["affine"] = bool(rng.random() < 0.5)
    else:
        specs = P.affine_specs(seed, n=int(rng.integers(1, 5)), fw=max(fw, 60), fh=max(fh, 50), homography=(kind == "homography"))
        if kind == "affine" and rng.random() < 0.6:
            # the global stage's device-side masks: black wedges in the strips, random low-resolution seam masks (or none),
            # random blur width, random gains
            sigma = float(rng.choice([10.0, 10.0, 3.0, 6.5, 1.2]))
            for s in specs:
                img = s["img"].copy()
                h_, w_ = img.shape[:2]
                yy, xx = np.mgrid[0:h_, 0:w_]
                img[yy < rng.uniform(0, 0.3) * xx - rng.uniform(0, 20)] = 0
                img[rng.integers(0, h_, 40), rng.integers(0, w_, 40)] = rng.integers(0, 6, (40, 3), dtype=np.uint8)
                s["img"] = img
                s["global_stage"] = True
                s["sigma"] = sigma
                if rng.random() < 0.7:
                    bw, bh = s["size"]
                    lw, lh = max(2, int(bw * rng.uniform(0.05, 0.5))), max(2, int(bh * rng.uniform(0.05, 0.5)))
                    m = (rng.random((lh, lw)) < 0.8).astype(np.uint8) * 255
                    m[rng.integers(0, lh), rng.integers(0, lw)] = 1
                    s["seam_lowres"] = m
                if rng.random() < 0.5:
                    s["gain"] = tuple(float(v) for v in rng.uniform(0.8, 1.25, 3))
            desc += f" global-stage sigma={sigma}"
    P.run_case(lib, specs, blend, bands, check_taps=bool(rng.random() < 0.5), band_split=split or None,
               out_format="bgra" if rng.random() < 0.2 else "bgr")
    if rng.random() < 0.5:
        desc += " crop:" + crop_case(lib, rng, seed)
    return desc


def crop_case(lib, rng, seed):
    """autoCropBlackBorder's rectangle on a small random mosaic: equal to the oracle's, or refused."""
    from drone_image_stitch_cpp_b200 import _lib as L, compositor as CP
    from oracle import ds_oracle as O
    sv = synth.grid_survey(int(rng.integers(1, 3)), int(rng.integers(1, 3)), int(rng.integers(40, 200)), int(rng.integers(40, 160)),
                           overlap=float(rng.uniform(0.2, 0.7)), seed=seed, rot_deg=float(rng.choice([0.0, 5.0, 30.0])))
    frames = [f.copy() for f in sv.frames]
    for f in frames:   # dark specks and stripes inside the content
        f[rng.integers(0, f.shape[0], 30), rng.integers(0, f.shape[1], 30)] = 0
        if rng.random() < 0.3:
            f[:, f.shape[1] // 2] = 0
    xfs = [CP.plane_transform(K, R, sv.scale) for K, R in zip(sv.Ks, sv.Rs)]
    rois = [CP.warp_roi(xf, f.shape[1], f.shape[0], lib) for xf, f in zip(xfs, frames)]
    cv = CP.Canvas(CP.result_roi(rois), "feather" if rng.random() < 0.5 else "multiband", 2, lib=lib)
    for i, (f, xf) in enumerate(zip(frames, xfs)):
        cv.upload(i, f, xf)
    cv.composite()
    pano, _ = cv.download()
    try:
        rect = cv.auto_crop_rect()
    except L.DroneStitchError as e:
        assert e.code == L.DS_ERR_UNSUPPORTED, e
        cv.close()
        return "refused"
    cv.close()
    assert rect == O.auto_crop_rect(pano), (rect, O.auto_crop_rect(pano))
    return "ok"
