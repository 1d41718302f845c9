"""Composite over resident frames (cfg2): monolithic schedule vs row slices of various heights (the level >= 1 feeds and
collapses of slice b overlap the level-0 feed of slice b+1 on a second stream). Wall clock over 20 queued composites.
usage: python tools/slice_sweep.py [rows ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drone_image_stitch_cpp_b200 import _lib, compositor as CP, synth
lib = _lib.default_library()
plan = synth.plan_grid(3, 3, 5472, 3648, overlap=0.7, seed=synth.MASTER_SEED)
xfs = [CP.plane_transform(K, R, plan.scale) for K, R in zip(plan.Ks, plan.Rs)]
rois = [CP.warp_roi(xf, plan.fw, plan.fh, lib) for xf in xfs]
roi = CP.result_roi(rois)
frames = synth.cut(plan, None, "cuda", as_torch=True)
torch.cuda.synchronize()
for rows in [int(v) for v in sys.argv[1:]] or [-1, 3072, 2048, 1536, 1024, 768, 512]:
    cv = CP.Canvas(roi, "multiband", 5, lib=lib, pipeline_rows=rows)
    for i, f in enumerate(frames):
        cv.upload_device(i, f.data_ptr(), plan.fw, plan.fh, plan.fw * 3, xfs[i])
    for _ in range(3):
        cv.composite_async()
    cv.synchronize()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(20):
            cv.composite_async()
        cv.synchronize()
        best = min(best, (time.perf_counter() - t0) / 20 * 1e3)
    pano, _ = cv.download()
    print(f"pipeline_rows {rows:5d}: {best:.3f} ms per composite, {roi[2] * roi[3] / 1e6 / best * 1e3:.0f} MP/s, checksum {int(pano[::7, ::7].sum())}", flush=True)
    cv.close()
