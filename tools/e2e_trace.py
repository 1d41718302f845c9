"""DS_TRACE=1 timeline of one pipelined end-to-end step (cfg2)."""
import os, sys, time, json
os.environ["DS_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drone_image_stitch_cpp_b200 import _lib, compositor as CP, synth
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 768
lib = _lib.default_library()
plan = synth.plan_grid(3, 3, 5472, 3648, overlap=0.7, seed=synth.MASTER_SEED)
xfs = [CP.plane_transform(K, R, plan.scale) for K, R in zip(plan.Ks, plan.Rs)]
rois = [CP.warp_roi(xf, plan.fw, plan.fh, lib) for xf in xfs]
roi = CP.result_roi(rois)
print("roi", roi, "frame rois", rois)
host = [torch.zeros((plan.fh, plan.fw, 3), dtype=torch.uint8).pin_memory() for _ in xfs]
out = torch.empty((roi[3], roi[2], 3), dtype=torch.uint8).pin_memory()
cv = CP.Canvas(roi, blend="multiband", bands=5, lib=lib, pipeline_rows=rows)
for rep in range(3):
    print("---- step", rep, file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    for i, h in enumerate(host):
        cv.upload(i, h.numpy(), xfs[i], async_=True)
    t1 = time.perf_counter()
    cv.composite_async()
    t2 = time.perf_counter()
    cv.download(out=out.numpy(), want_mask=False)
    t3 = time.perf_counter()
    cv.synchronize()
    print(f"host: uploads queued {1e3*(t1-t0):.3f} ms, composite queued {1e3*(t2-t1):.3f} ms, download returned {1e3*(t3-t2):.3f} ms", file=sys.stderr, flush=True)
