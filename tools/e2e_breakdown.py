"""Where one end-to-end step (upload from pinned host + composite + download) spends its time."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drone_image_stitch_cpp_b200 import _lib, compositor as CP, synth

lib = _lib.default_library()
plan = synth.plan_grid(3, 3, 5472, 3648, overlap=0.7, seed=synth.MASTER_SEED)
xfs = [CP.plane_transform(K, R, plan.scale) for K, R in zip(plan.Ks, plan.Rs)]
rois = [CP.warp_roi(xf, plan.fw, plan.fh, lib) for xf in xfs]
roi = CP.result_roi(rois)
rng = np.random.default_rng(1)
host = []
for i in range(len(xfs)):
    t = torch.empty((plan.fh, plan.fw, 3), dtype=torch.uint8).pin_memory()
    t.numpy()[:] = rng.integers(0, 255, (plan.fh, plan.fw, 3), dtype=np.uint8)
    host.append(t)
out = torch.empty((roi[3], roi[2], 3), dtype=torch.uint8).pin_memory()
cv = CP.Canvas(roi, blend="multiband", bands=5, lib=lib)
res = {}
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i, h in enumerate(host):
        cv.upload(i, h.numpy(), xfs[i])
    t1 = time.perf_counter()
    cv.composite()
    t2 = time.perf_counter()
    cv.download(out=out.numpy(), want_mask=False)
    t3 = time.perf_counter()
    res = {"upload_ms": (t1 - t0) * 1e3, "composite_ms": (t2 - t1) * 1e3, "download_ms": (t3 - t2) * 1e3,
           "h2d_GBps": sum(h.numel() for h in host) / (t1 - t0) / 1e9, "d2h_GBps": out.numel() / (t3 - t2) / 1e9,
           "kernel_ms": cv.info().ms_last_composite}
    print(json.dumps(res), flush=True)
cv.close()
# pipelined schedule: asynchronous uploads, composite in row slices, download slice by slice
for rows in [int(v) for v in os.environ.get("SWEEP_ROWS", "-1,256,384,512,768,1024,1536,2048,3072").split(",")]:
    cv = CP.Canvas(roi, blend="multiband", bands=5, lib=lib, pipeline_rows=rows)
    def step():
        for i, h in enumerate(host):
            cv.upload(i, h.numpy(), xfs[i], async_=True)
        cv.composite_async()
        cv.download(out=out.numpy(), want_mask=False)
        cv.synchronize()
    step(); step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    dt = (time.perf_counter() - t0) / 5
    print(json.dumps({"pipeline_rows": rows, "e2e_ms": dt * 1e3, "MPps": roi[2] * roi[3] / 1e6 / dt,
                      "launches": int(cv.info().launches_last_composite)}), flush=True)
    if "ref" not in globals():
        ref = out.numpy().copy()
    assert np.array_equal(ref, out.numpy()), "pipelined result differs"
    cv.close()
# raw PCIe rates for comparison
d = torch.empty(host[0].numel(), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for h in host: d.copy_(h.view(-1), non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(json.dumps({"raw_h2d_GBps": sum(h.numel() for h in host) / (t1 - t0) / 1e9}))
dd = torch.empty(out.numel(), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
out.view(-1).copy_(dd, non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(json.dumps({"raw_d2h_GBps": out.numel() / (t1 - t0) / 1e9}))
# both directions at once
s2 = torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s2):
    for _ in range(3): out.view(-1).copy_(dd, non_blocking=True)
for h in host: d.copy_(h.view(-1), non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(json.dumps({"duplex_ms": (t1 - t0) * 1e3, "duplex_GBps_sum": (sum(h.numel() for h in host) + 3 * out.numel()) / (t1 - t0) / 1e9}))
