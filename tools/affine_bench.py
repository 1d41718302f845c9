"""cfg2 frames in the forms the reference's two call sites really use, to time the level-0 code paths they take:
  affine      AFFINE_F64 transforms (the global stage's cv::warpAffine form, stitch_global.cpp:474-480)
  affine_seam the same with a soft blend mask and a channel gain per strip (stitch_global.cpp:644-658)
  plane_seam  PLANE_F32 with a seam mask and a block gain map per frame (composePanorama with DpSeamFinder and
              BlocksGainCompensator, stitch_robust.cpp:207-211)
  plane_proj  PLANE_F32 with a slight perspective (the map divides by z per pixel: per-pixel loop of the fast kernel)
  homography  HOMOGRAPHY_F64 transforms (cv::warpPerspective coordinates, double arithmetic per pixel): generic level-0 kernel
  many        30 PLANE_F32 frames over the same ground (more than the 24 frames per tile the fast level-0 kernel stages): generic kernel
usage: python tools/affine_bench.py [mode]"""
import os, sys, json, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drone_image_stitch_cpp_b200 import _lib, compositor as CP, synth

mode = sys.argv[1] if len(sys.argv) > 1 else "affine"
lib = _lib.default_library()
plan = synth.plan_grid(3, 3, 5472, 3648, overlap=0.7, seed=synth.MASTER_SEED)
fw, fh = plan.fw, plan.fh
xfs, rois = [], []
for A in plan.A:
    # transformedBoundingRect (stitch_global.cpp:71-98): corners (0,0),(w,0),(w,h),(0,h) in double, floor(min), ceil(max) - x
    pts = np.array([[0, 0, 1], [fw, 0, 1], [fw, fh, 1], [0, fh, 1]], np.float64) @ A.T
    x0, y0 = math.floor(pts[:, 0].min()), math.floor(pts[:, 1].min())
    w, h = max(1, math.ceil(pts[:, 0].max()) - x0), max(1, math.ceil(pts[:, 1].max()) - y0)
    M = A.copy(); M[0, 2] -= x0; M[1, 2] -= y0
    xfs.append(CP.affine_transform(M, (x0, y0), (w, h)))
    rois.append((x0, y0, w, h))
if mode == "plane_proj":
    # PlaneWarper with a slight perspective (mode C of SURVEY 8(d)): the map divides by z per pixel
    Rs = []
    for R in plan.Rs:
        R = np.array(R, np.float32).copy(); R[2, 0] = 1e-6; R[2, 1] = -1e-6
        Rs.append(R)
    xfs = [CP.plane_transform(K, R, plan.scale, affine=False) for K, R in zip(plan.Ks, Rs)]
    rois = [CP.warp_roi(xf, fw, fh, lib) for xf in xfs]
if mode == "homography":
    xfs = []
    for A, r in zip(plan.A, rois):
        H = np.eye(3); H[:2] = A; H[0, 2] -= r[0]; H[1, 2] -= r[1]
        H[2, 0] = 1e-7; H[2, 1] = -1e-7
        xfs.append(CP.homography_transform(H, (r[0], r[1]), (r[2], r[3])))
if mode == "many":
    plan = synth.plan_grid(6, 5, 5472, 3648, overlap=0.93, seed=synth.MASTER_SEED)
    xfs = [CP.plane_transform(K, R, plan.scale) for K, R in zip(plan.Ks, plan.Rs)]
    rois = [CP.warp_roi(xf, fw, fh, lib) for xf in xfs]
if mode == "plane_seam":
    xfs = [CP.plane_transform(K, R, plan.scale) for K, R in zip(plan.Ks, plan.Rs)]
    rois = [CP.warp_roi(xf, fw, fh, lib) for xf in xfs]
roi = CP.result_roi(rois)
frames = synth.cut(plan, None, device="cpu")
cv = CP.Canvas(roi, "multiband", 5, lib=lib)
rng = np.random.default_rng(3)
for i, f in enumerate(frames):
    kw = {}
    if mode not in ("affine", "plane_proj", "homography", "many"):
        # a seam mask over the frame's bbox: everything but a diagonal band and the outer 40 px; soft edges for the global stage
        w, h = rois[i][2], rois[i][3]
        yy, xx = np.mgrid[0:h, 0:w]
        m = np.full((h, w), 255, np.uint8)
        m[np.abs(xx * h - yy * w) < 0.02 * w * h] = 0
        m[:40] = 0; m[-40:] = 0; m[:, :40] = 0; m[:, -40:] = 0
        if mode == "affine_seam":
            m[(np.abs(xx * h - yy * w) >= 0.02 * w * h) & (np.abs(xx * h - yy * w) < 0.03 * w * h)] = 128
            kw = dict(seam_mask=m, channel_gain=[1.02, 0.99, 1.01])
        else:
            kw = dict(seam_mask=m, gain_map=(1.0 + 0.05 * rng.random((h, w))).astype(np.float32))
    cv.upload(i, f, xfs[i], **kw)
for _ in range(3):
    cv.composite()
cv.set_profiling(True)
for _ in range(5):
    cv.composite()
kt = cv.kernel_times()
agg = {}
for k in kt:
    agg.setdefault((k["name"], k["level"]), []).append(k["ms"])
print(json.dumps({"mode": mode, "canvas": [roi[2], roi[3]], "frames": len(frames), "MP_per_s": roi[2] * roi[3] / 1e3 / cv.info().ms_last_composite, "ms_composite": cv.info().ms_last_composite,
                  "kernels": {f"{n}[{l}]": round(float(np.mean(v)), 4) for (n, l), v in agg.items()}}))
