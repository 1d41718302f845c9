"""BASELINE config 4 scale on ONE B200: 600 frames of 5472x3648 (12 lines x 50, 70 % forward / 32 % side overlap),
multi-band 5, canvas ~2.5 GP. Frames are generated on the device (seeded noise: only indexing and memory matter
here) and handed over with ds_upload_frame_device. Checks: the composite runs, is idempotent, and three row bands
computed by separate band handles reproduce the same rows of the big canvas bit for bit (64-bit indexing, tile
lists and TMA descriptors at > 2^31 pixels). Prints one JSON line.
With a third argument (path), a 2048x1536 window from the middle of the canvas is written there as .npz (tile, mask, win,
roi) - tests/test_gpu_parity.py compares it with the windowed oracle.
usage: python tools/scale_check.py [nx ny [window.npz]]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drone_image_stitch_cpp_b200 import _lib, compositor as CP, synth

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (50, 12)
fw, fh, bands = 5472, 3648, 5
lib = _lib.default_library()
plan = synth.plan_grid(nx, ny, fw, fh, overlap=0.7, side_overlap=0.32, seed=synth.MASTER_SEED)
xfs = [CP.plane_transform(K, R, plan.scale) for K, R in zip(plan.Ks, plan.Rs)]
rois = [CP.warp_roi(xf, fw, fh, lib) for xf in xfs]
roi = CP.result_roi(rois)
print("canvas", roi[2], "x", roi[3], "=", roi[2] * roi[3] / 1e9, "GP;", len(xfs), "frames", flush=True)


def frame(i):
    g = torch.Generator(device="cuda").manual_seed(1000 + i)
    f = torch.randint(0, 256, (fh, fw, 3), dtype=torch.uint8, device="cuda", generator=g)
    torch.cuda.synchronize()   # the library reads the pixels on its own streams
    return f


def fill(cv, which):
    for i in which:
        f = frame(i)
        cv.upload_device(i, f.data_ptr(), fw, fh, fw * 3, xfs[i])
        del f


t0 = time.time()
cv = CP.Canvas(roi, "multiband", bands, lib=lib)
fill(cv, range(len(xfs)))
t_up = time.time() - t0
cv.composite()
ms = []
for _ in range(3):
    cv.composite()
    ms.append(cv.info().ms_last_composite)
info = cv.info()
H, m = info.padded_height, 1 << info.num_bands
res = {"canvas": [roi[2], roi[3]], "gigapixels": roi[2] * roi[3] / 1e9, "frames": len(xfs), "device_GB": info.device_bytes / 1e9,
       "composite_ms": min(ms), "MPps": roi[2] * roi[3] / 1e6 / (min(ms) / 1e3), "upload_s": t_up, "bands_checked": []}
print(json.dumps(res), flush=True)
if len(sys.argv) > 3:
    ww, wh = 2048, 1536
    wx, wy = (roi[2] - ww) // 2 // m * m, (roi[3] - wh) // 2 // m * m
    tile, tmask = cv.download(wx, wy, ww, wh)
    np.savez(sys.argv[3], tile=tile, mask=tmask, win=np.array([wx, wy, ww, wh]), roi=np.array(roi))
# three row bands of 512 rows: top, across the middle, bottom
for y0 in (0, (H // 2) // m * m, (roi[3] - 512) // m * m):
    y1 = min(y0 + 512, H)
    big, bigmask = cv.download(y=y0, h=min(y1, roi[3]) - y0)
    cb = CP.Canvas(roi, "multiband", bands, band=(y0, y1), lib=lib)
    fill(cb, [i for i in range(len(xfs)) if cb.touches(rois[i])])
    cb.composite()
    rows, rmask = cb.download()
    ok = bool(np.array_equal(rows, big) and np.array_equal(rmask, bigmask))
    res["bands_checked"].append({"rows": [y0, y1], "frames": int(cb.info().num_frames), "identical": ok, "nonzero": int(np.count_nonzero(bigmask))})
    cb.close()
    del big, rows
res["ok"] = all(b["identical"] and b["nonzero"] > 0 for b in res["bands_checked"])
print(json.dumps(res), flush=True)
cv.close()
sys.exit(0 if res["ok"] else 1)
