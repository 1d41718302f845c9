#!/usr/bin/env python
"""bench.py — output canvas megapixels/sec of the compositing hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (OpenCV), rank 0

Workload (default): BASELINE configs[3] — ONE 600-frame survey (12 lines x 50 frames of 5472x3648, 70 % forward /
32 % side overlap, multi-band 5) composited into one 2.7-gigapixel canvas. It is the largest configuration that fits
one B200 (125 GB of HBM); at N > 1 the SAME canvas is cut into N row bands, one per GPU / process (strong scaling),
each band holding only the frames that touch it, neighbouring bands handing each other their level-1 pyramid halo
rows over NVLink P2P (ds_p2p_connect; no collective on the data path — torch.distributed carries the IPC handles
once, the barrier, and the max-over-ranks of the device time). `--workload cfg2|cfg1|cfg3|small` select the other
BASELINE configurations; at N = 1 the line also carries `also.cfg2` / `also.cfg1` for continuity with round 1.

A "step" is one ds_composite over the frames resident in HBM (`value`, CUDA events), or the upload of every frame
from pinned host memory + composite + download of the panorama through the C ABI (`e2e`).
Every run checks its own output: each rank compares windows at its band edges with the windowed CPU oracle
(oracle/windowed.py, checker only) and the line carries `parity`; a mismatch exits non-zero.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "output_canvas_megapixels_per_sec"
UNIT = "MP/s"
DTYPE = "u8/int16/f32"


# ------------------------------------------------------------------------------------------------ workload

def survey(name):
    from drone_image_stitch_cpp_b200 import synth
    return synth.plan_survey(name)


def geometry(plan, lib):
    from drone_image_stitch_cpp_b200 import compositor as CP
    xfs = [CP.plane_transform(K, R, plan.scale) for K, R in zip(plan.Ks, plan.Rs)]
    rois = [CP.warp_roi(xf, plan.fw, plan.fh, lib) for xf in xfs]
    return xfs, rois, CP.result_roi(rois)


def job_config(desc, plan, roi, blend, bands, n_gpus):
    """The `config` object of the JSON line: a function of the workload and N only, so both arms print the same one."""
    frame_bytes = plan.fw * plan.fh * 4 * len(plan.A)
    return {"workload": desc, "canvas": [int(roi[2]), int(roi[3])], "frames": len(plan.A), "frame_size": [plan.fw, plan.fh],
            "blend": blend, "bands": int(bands), "parallelism": f"row-bands x{n_gpus} of one canvas",
            "l2": (f"inputs larger than L2 ({frame_bytes / n_gpus / 1e6:.0f} MB of frames per GPU vs 126 MB)" if frame_bytes / n_gpus > 3 * 126e6
                   else f"L2 flushed between timed composites (a 256 MB buffer is rewritten; the {frame_bytes / n_gpus / 1e6:.0f} MB of frames would fit L2)")}


def make_frame(plan, i, device, procedural, ortho_cache):
    """Frame i as an HxWx3 uint8 torch tensor on `device`."""
    from drone_image_stitch_cpp_b200 import synth
    if procedural:
        return synth.procedural_frame(plan, i, device=device, as_torch=True)
    if "o" not in ortho_cache:
        ortho_cache["o"] = synth.orthophoto(plan.ortho_h, plan.ortho_w, plan.seed, device)
    return synth.cut(plan, [i], device=device, as_torch=True, ortho=ortho_cache["o"])[0]


def uses_procedural(plan):
    # the stored orthophoto (float CHW on the device) is only practical for small canvases
    return plan.ortho_w * plan.ortho_h > 400e6


# ------------------------------------------------------------------------------------------------ clocks

class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled every ~2 ms from a thread
    (nvidia-smi -lms is the fallback; its start-up alone is longer than a short timed region)."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, index):
        self.index, self.sm, self.mask, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self.t = None
        self.src = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[0].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.src = "nvml"
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.src = None

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self.src != "nvml":
            return self._smi_once()
        self._stop.set()
        self.t.join(timeout=1.0)
        reasons = sorted(n for n, bit in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.sm), "source": "nvml polled every 2 ms during the timed region"}

    def _smi_once(self):
        try:
            out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                           "--format=csv,noheader,nounits"], text=True, timeout=10)
            f = [x.strip() for x in out.strip().split(",")]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "reasons": [], "samples": 1,
                    "source": "nvidia-smi once after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def dominant_traffic(workload, name, level):
    """ncu DRAM bytes per launch of the dominant kernel, from the committed capture summary - only if that capture
    was taken from the kernels this library was built from (source hash recorded beside it)."""
    tp = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    try:
        from drone_image_stitch_cpp_b200 import build
        d = json.load(open(tp))
        if d.get("kernel_source_sha16") != build.source_hash():
            return None, f"capture is of other kernel sources ({d.get('kernel_source_sha16')}), not reported"
        return d.get(f"{workload}:{name}:{level}"), d.get("capture")
    except Exception:
        return None, "no capture summary"


# ------------------------------------------------------------------------------------------------ CPU reference

class CpuReference:
    """The reference's CPU path (OpenCV through oracle/cv_reference.py, else the C port) on a bounded sample of the
    workload at FULL frame size: the first `cols` frames of every flight line (a shorter flight of the same pattern:
    same overlaps, same frame size, same blend), composited whole. Synthetic frames are generated once, outside the
    timed region (on the GPU when one is present)."""

    def __init__(self, plan, blend, bands, nx, target_seconds, rate_guess=11.0, threads=None):
        from oracle import cv_reference as CR
        from oracle import ds_oracle as O
        self.use_cv = CR.have_cv2()
        cores = os.cpu_count() or 1
        if self.use_cv:
            self.cores = CR.set_threads(threads or cores)
        else:
            O.set_threads(threads or cores)
            self.cores = O.get_threads()
        self.kind = "reference" if self.use_cv else "port"
        self.impl = ("OpenCV (cv2 wheel): AffineWarper + MultiBandBlender/FeatherBlender" if self.use_cv else "oracle/ds_oracle.c (OpenMP)")
        self.blend, self.bands, self.plan = blend, bands, plan
        n = len(plan.A)
        ny = n // nx
        # columns so that one compose takes about target_seconds at rate_guess output MP/s
        stepx = plan.fw * 0.3
        canvas_h = plan.fh * (1 + 0.68 * (ny - 1)) if ny > 1 else plan.fh
        want_w = target_seconds * rate_guess * 1e6 / canvas_h
        cols = int(max(1, min(nx, round((want_w - plan.fw) / stepx) + 1)))
        self.cols, self.ny, self.nx = cols, ny, nx
        self.idx = []
        for jj in range(ny):
            for i in range(nx):
                col = i if jj % 2 == 0 else nx - 1 - i    # serpentine flight order (synth.plan_grid)
                if col < cols:
                    self.idx.append(jj * nx + i)
        import torch
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        proc = uses_procedural(plan)
        cache = {}
        self.frames = [make_frame(plan, i, dev, proc, cache).cpu().numpy() for i in self.idx]
        cache.clear()
        self.Ks = [plan.Ks[i] for i in self.idx]
        self.Rs = [plan.Rs[i] for i in self.idx]
        self.last_dt = None

    def step(self):
        from oracle import cv_reference as CR
        from oracle import ds_oracle as O
        t0 = time.perf_counter()
        if self.use_cv:
            pano, _, roi = CR.compose_cv2(self.frames, self.Ks, self.Rs, self.plan.scale, self.blend, self.bands)
        else:
            pano, _, roi = O.compose_port(self.frames, self.Ks, self.Rs, self.plan.scale, self.blend, self.bands)
        dt = time.perf_counter() - t0
        self.last_dt, self.last_roi = dt, roi
        return roi[2] * roi[3] / 1e6, dt

    def sample(self):
        whole = self.cols == self.nx
        what = ("the whole workload" if whole else
                f"the first {self.cols} frames of each of the {self.ny} flight lines ({len(self.idx)} of {len(self.plan.A)} frames)")
        return (f"{what}, full-size {self.plan.fw}x{self.plan.fh} frames, canvas {self.last_roi[2]}x{self.last_roi[3]} = "
                f"{self.last_roi[2] * self.last_roi[3] / 1e6:.0f} MP, {self.impl}, one compose = {self.last_dt:.2f} s")


def grid_nx(name):
    return {"cfg4": 50, "cfg3": 40, "cfg5": 80, "cfg2": 3, "cfg1": 2, "small": 6}[name]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from drone_image_stitch_cpp_b200 import _lib
    lib = _lib.default_library()
    plan, blend, bands, desc = survey(args.workload)
    xfs, rois, roi = geometry(plan, lib)
    total = args.steps + args.warmup
    # one compose of the sample about 45 s (cfg4: 12 lines x 7 frames, 0.5 GP) unless the whole workload is smaller
    ref = CpuReference(plan, blend, bands, grid_nx(args.workload), target_seconds=45.0)
    mps, secs, nsteps = 0.0, 0.0, 0
    t_all = time.perf_counter()
    budget = 200.0
    warm = min(args.warmup, 1) if ref.cols < ref.nx or plan.fw * plan.fh * len(plan.A) > 100e6 else args.warmup
    for i in range(total):
        mp, dt = ref.step()
        if i >= warm:
            mps += mp; secs += dt; nsteps += 1
        if nsteps >= args.steps:
            break
        if nsteps >= 1 and time.perf_counter() - t_all + dt > budget:
            break
    v = mps / secs
    sample = ref.sample() + f"; {nsteps} timed compose(s) after {warm} warm-up (bounded to {budget:.0f} s)"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": nsteps,
            "warmup": warm, "ms_per_step": secs / nsteps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic",
            "config": job_config(desc, plan, roi, blend, bands, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ref.cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ native arm

def bind_to_gpu_numa_node(idx):
    """Pin this process (and so its pinned host buffers, first-touch) to the CPU cores NVML reports as local to
    GPU `idx`: with several ranks per box the host<->device copies otherwise cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        node = None
        try:
            node = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
        if node is None:
            try:   # sysfs knows the node of the GPU's PCI function even where NVML does not report it
                bus = pynvml.nvmlDeviceGetPciInfo(h).busId
                bus = bus.decode() if isinstance(bus, bytes) else bus
                with open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node") as fh:
                    node = int(fh.read().strip())
            except Exception:
                pass
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"gpu": idx, "cores": len(cpus), "first_core": min(cpus), "numa_node": node}
    except Exception as e:  # measurement nicety only
        return {"gpu": idx, "unbound": type(e).__name__}
    return {"gpu": idx, "unbound": "no affinity reported"}


class L2Flusher:
    """Rewrites a buffer twice the size of L2 between timed composites (for workloads whose frames would fit L2)."""

    def __init__(self, device, stream):
        import torch
        self.buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
        self.stream = stream

    def __call__(self):
        import torch
        with torch.cuda.stream(self.stream):
            self.buf.add_(1)


def resident_value(cv, steps, warmup, stream, barrier, flush=None):
    """K timed composites over frames resident in HBM: device time by CUDA events on the canvas stream."""
    import torch
    for _ in range(max(warmup, 3)):
        cv.composite_async()
    cv.set_profiling(True)
    barrier()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for k in range(steps):
        if flush is not None:
            flush()
        e0[k].record(stream)
        cv.composite_async()
        e1[k].record(stream)
    barrier()
    ms = [a.elapsed_time(b) for a, b in zip(e0, e1)]
    if flush is None:
        total = e0[0].elapsed_time(e1[-1])     # back to back: one span
    else:
        total = float(sum(ms))                  # flushes between the composites are not part of the steps
    kt = cv.kernel_times()
    cv.set_profiling(False)
    return total, kt


def quick_workload(name, lib, local, steps=10):
    """A small single-GPU side measurement for the `also` object (composite-only, frames resident)."""
    import torch
    from drone_image_stitch_cpp_b200 import compositor as CP
    plan, blend, bands, desc = survey(name)
    xfs, rois, roi = geometry(plan, lib)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        cv = CP.Canvas(roi, blend, bands, out_format="bgr", device=local, stream=stream.cuda_stream, lib=lib)
        cache = {}
        for i in range(len(xfs)):
            f = make_frame(plan, i, f"cuda:{local}", uses_procedural(plan), cache)
            torch.cuda.synchronize()
            cv.upload_device(i, f.data_ptr(), plan.fw, plan.fh, plan.fw * 3, xfs[i])
            del f
        cache.clear()
        flush = None
        if plan.fw * plan.fh * 4 * len(xfs) <= 3 * 126e6:
            flush = L2Flusher(f"cuda:{local}", stream)

        def barrier():
            torch.cuda.synchronize()
        total, kt = resident_value(cv, steps, 3, stream, barrier, flush)
        ab = int(cv.info().algorithmic_bytes)
        cv.close()
    mp = roi[2] * roi[3] / 1e6
    ms = total / steps
    peak, _ = hbm_peak()
    agg = {}
    for k in kt:
        agg.setdefault((k["name"], k["level"]), []).append(k["ms"])
    top = max(agg.items(), key=lambda kv: np.mean(kv[1]))
    top_ab = next(k["algorithmic_bytes"] for k in kt if (k["name"], k["level"]) == top[0])
    top_ms = float(np.mean(top[1]))
    return {"workload": desc, "canvas": [int(roi[2]), int(roi[3])], "value": mp / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
            "steps": steps, "l2": "flushed between composites" if flush is not None else "inputs larger than L2",
            "whole_step_frac": ab / (ms / 1e3) / 1e9 / peak,
            "dominant": {"kernel": f"{top[0][0]}[level {top[0][1]}]", "ms_per_launch": top_ms,
                         "frac": top_ab / (top_ms / 1e3) / 1e9 / peak}}


def parity_windows(edges, rank, world, roi, bands, PH):
    """Windows (x, y, w, h) this rank checks and the canvas rows of its own it compares in each: one window straddling
    every band edge the rank touches (compared on its side of the edge), or the canvas centre on a single GPU."""
    m = 1 << bands
    g = 8 << bands                      # margin of the windowed oracle: 8 * 2^L px inside the window
    ww = min(max(2048, 2 * g + 1024), (roi[2] // m) * m)
    wx = max(0, ((roi[2] - ww) // 2) // m * m)
    reach = max(768, g + 256)           # rows either side of an edge: the margin + at least 256 compared rows
    half = -(-reach // m) * m if PH >= 4 * reach else max(g + m, (PH // 4) // m * m)
    out = []

    def win_at(e):
        y0 = max(0, e - half)
        y1 = min(PH, e + half)
        return (wx, y0, ww, y1 - y0)
    if world == 1:
        out.append((win_at((PH // 2) // m * m), None))
    elif bands >= 7:
        # deep pyramids: one 2 g + 1024 wide window costs the CPU oracle about a minute, so two ranks check one edge each
        # (the band below the first edge and the one below the middle edge: both sides of a P2P hand-over are exercised
        # by the pull of the rank that checks); the other ranks' rows are covered by tests/test_gpu_parity.py (cfg5 band)
        if rank in (1, world // 2):
            out.append((win_at(edges[rank]), "below"))
    else:
        if rank > 0:
            out.append((win_at(edges[rank]), "below"))          # my rows just below my upper edge
        if rank < world - 1:
            out.append((win_at(edges[rank + 1]), "above"))      # my rows just above my lower edge
    return out, g


def run_native(args):
    import torch
    import torch.distributed as dist
    from drone_image_stitch_cpp_b200 import _lib, compositor as CP

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (libdronestitch_cuda has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.default_library()
    t_start = time.perf_counter()

    def note(msg):
        if args.verbose and rank == 0:
            print(f"[bench {time.perf_counter() - t_start:7.1f}s] {msg}", file=sys.stderr, flush=True)

    # ---------------- side measurements for continuity (single GPU only, before the big canvas takes the memory)
    also = {}
    if world == 1 and not args.no_also and args.workload == "cfg4":
        for name in ("cfg2", "cfg1"):
            try:
                also[name] = quick_workload(name, lib, local)
            except Exception as e:   # never lose the headline over a side measurement
                also[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
            note(f"also.{name} done")
        torch.cuda.empty_cache()

    plan, blend, bands, desc = survey(args.workload)
    if args.bands is not None and blend == "multiband":
        bands = args.bands
        desc += f" [bands overridden: {bands}]"
    xfs, rois, roi = geometry(plan, lib)
    procedural = uses_procedural(plan)
    probe = CP.Canvas(roi, blend, bands, lib=lib, device=local)
    pinfo = probe.info()
    probe.close()
    PH, eff_bands = pinfo.padded_height, pinfo.num_bands
    edges = CP.plan_row_bands(roi, rois, world, blend, bands, lib=lib) if world > 1 else [0, PH]
    band = (edges[rank], edges[rank + 1]) if world > 1 else None

    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        cv = CP.Canvas(roi, blend, bands, out_format="bgr", device=local, band=band, stream=stream.cuda_stream, lib=lib)
        mine = [i for i in range(len(xfs)) if band is None or cv.touches(rois[i])]
        frame_bytes = plan.fw * plan.fh * 3
        # pinned host copies of this rank's frames (the e2e step uploads them); one allocation, first touched here
        host_mode = "every frame of the rank in pinned host memory"
        pool = None
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 1 << 62
        n_host = len(mine)
        if n_host * frame_bytes > 0.6 * avail / world or args.host_pool:   # (every rank of the box pins its share)
            n_host = max(8, min(len(mine), int(args.host_pool or 64)))
            host_mode = f"a pool of {n_host} pinned frame buffers cycled over the rank's {len(mine)} frames (host memory)"
        pinned = torch.empty((n_host, plan.fh, plan.fw, 3), dtype=torch.uint8, pin_memory=True)
        cache = {}
        for k, i in enumerate(mine):
            f = make_frame(plan, i, dev, procedural, cache)
            torch.cuda.synchronize()
            cv.upload_device(i, f.data_ptr(), plan.fw, plan.fh, plan.fw * 3, xfs[i])
            if k < n_host:
                pinned[k].copy_(f)
            torch.cuda.synchronize()
            del f
        cache.clear()
        host_np = [pinned[k % n_host].numpy() for k in range(len(mine))]
        info = cv.info()
        out_rows = min(info.band_y1, roi[3]) - info.band_y0
        out_pin = torch.empty((max(out_rows, 1), roi[2], 3), dtype=torch.uint8, pin_memory=True)
        note(f"{len(mine)} frames resident, {info.device_bytes / 1e9:.1f} GB of HBM")

        # neighbouring bands hand each other the level-1 halo rows over NVLink (no collective on the data path;
        # torch.distributed only carries the IPC handles once, here)
        halo = "none (one band)" if world == 1 else "recomputed per band (no exchange)"
        use_p2p = world > 1 and blend == "multiband" and args.p2p != "off"
        if use_p2p:
            blobs = [None] * world
            dist.all_gather_object(blobs, cv.p2p_export())
            ok, why = True, ""
            try:
                if rank > 0:
                    cv.p2p_connect(0, blobs[rank - 1])
                if rank < world - 1:
                    cv.p2p_connect(1, blobs[rank + 1])
            except _lib.DroneStitchError as e:
                ok, why = False, str(e)
            oks = [None] * world
            dist.all_gather_object(oks, (ok, why))
            if all(o[0] for o in oks):
                halo = "NVLink P2P pull of the level-1 halo rows (ds_p2p_connect); recomputed in the pipelined e2e schedule"
            else:
                cv.p2p_disconnect()
                halo = "recomputed per band (P2P unavailable: " + next(o[1] for o in oks if not o[0])[:120] + ")"

        def gather_rows(rows):
            if world == 1:
                return [rows]
            out = [None] * world
            dist.all_gather_object(out, rows)
            return out

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def allmax(x):
            t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        def allsum(xs):
            t = torch.tensor([float(v) for v in xs], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t)
            return [float(v) for v in t.tolist()]

        # ---------------- value: composite-only, frames resident in HBM
        flush = None
        if plan.fw * plan.fh * 4 * len(xfs) / world <= 3 * 126e6:
            flush = L2Flusher(dev, stream)
        clocks = ClockSampler(local)
        clocks.start()
        ms_total, kt = resident_value(cv, args.steps, args.warmup, stream, barrier, flush)
        clk = clocks.stop()
        launches = int(cv.info().launches_last_composite) * args.steps
        ms_rank = gather_rows(round(ms_total / args.steps, 4))   # every rank's own device time per composite
        ms_total = allmax(ms_total)
        canvas_mp = roi[2] * roi[3] / 1e6
        value = canvas_mp * args.steps / (ms_total / 1e3)
        ab_total = int(cv.info().algorithmic_bytes)
        note(f"value {value:.0f} MP/s ({ms_total / args.steps:.2f} ms per composite)")

        # ---------------- parity: windows at this rank's band edges against the windowed CPU oracle (checker only)
        parity = {"checked": False}
        if not args.no_parity and blend == "multiband":
            from oracle import windowed as W
            wins, g = parity_windows(edges, rank, world, roi, eff_bands, PH)
            if world > 1:
                # (torchrun pins OMP_NUM_THREADS to 1; the checker may use this rank's share of the host cores)
                from oracle import ds_oracle as O
                checkers = 2 if eff_bands >= 7 else world
                O.set_threads(max(1, (os.cpu_count() or 1) // checkers))
            n_px = n_bad = mx = 0
            y_lo, y_hi = info.band_y0, min(info.band_y1, roi[3])
            checked_rows = []
            for win, side in wins:
                need = W.frames_touching(rois, roi, win)
                frames = {}
                for i in need:
                    if i in mine and mine.index(i) < n_host:
                        frames[i] = host_np[mine.index(i)]
                    else:   # a frame of the neighbouring band (or beyond the host pool): generated again, same function
                        frames[i] = make_frame(plan, i, dev, procedural, cache).cpu().numpy()
                cache.clear()
                ref, refmask = W.compose_window(frames, plan.Ks, plan.Rs, plan.scale, eff_bands, roi, win)
                ya, yb = max(win[1] + g, y_lo), min(win[1] + win[3] - g, y_hi)
                if ya >= yb:
                    continue
                rows, rmask = cv.download(x=win[0], y=ya, w=win[2], h=yb - ya)
                a, b, c = W.compare_inside(rows, rmask, ya, ref, refmask, win, eff_bands, x0=win[0])
                n_px += a; n_bad += b; mx = max(mx, c)
                checked_rows.append([int(ya), int(yb)])
            tot = allsum([n_px, n_bad])
            parity = {"checked": tot[0] > 0, "identical": tot[0] > 0 and tot[1] == 0, "pixels_compared": int(tot[0]),
                      "pixels_differing": int(tot[1]), "max_abs_diff": int(allmax(mx)),
                      "how": (("two ranks compare one" if (world > 1 and eff_bands >= 7) else "every rank compares") +
                              f" {max(2048, 2 * g + 1024)}-px-wide windows straddling band edges (the rank's own rows, >= 8 * 2^L = {g} px inside the "
                              "window) with the windowed CPU oracle (oracle/windowed.py, SURVEY 8(c) P17); bit-exact required"),
                      "rows_checked": [r for rr in (gather_rows(checked_rows)) for r in rr]}
            note(f"parity {parity['identical']} over {parity['pixels_compared']} px")

        # ---------------- e2e: upload from pinned host + composite + download, every step
        e2e_steps = max(2, min(args.steps, 5 if canvas_mp < 500 else 3))
        d2h = int(out_rows) * int(roi[2]) * 3

        def e2e_step():
            # the caller's frames stay valid for the step, so the uploads are queued (DS_UPLOAD_ASYNC), the composite
            # works through the canvas in row slices as the frames arrive and the download copies every slice out
            # as soon as it is final: host->device copies, kernels and device->host copies overlap
            for i, arr in zip(mine, host_np):
                cv.upload(i, arr, xfs[i], async_=True)
            cv.composite_async()
            cv.download(out=out_pin.numpy()[:max(out_rows, 0)], want_mask=False)
            cv.synchronize()

        e2e_step()
        barrier()
        h2d0 = int(cv.info().h2d_bytes_total)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = allmax(time.perf_counter() - t0)
        # bytes the library actually copied (a row-band handle only pulls the source rows its band reads)
        h2d = (int(cv.info().h2d_bytes_total) - h2d0) // e2e_steps
        e2e_val = canvas_mp * e2e_steps / dt
        e2e_ms = dt / e2e_steps * 1e3
        h2d_all, d2h_all = allsum([h2d, d2h])
        note(f"e2e {e2e_val:.0f} MP/s ({e2e_ms:.1f} ms per step)")

        # ---------------- the PCIe floor of that step: the same bytes, N ranks at once, pinned memory, no kernels
        floor_ms = None
        try:
            s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
            scratch = torch.empty((plan.fh, plan.fw, 3), dtype=torch.uint8, device=dev)
            dev_out = torch.empty_like(out_pin, device=dev)
            n_up = max(1, int(round(h2d / frame_bytes)))
            rem = h2d - (n_up - 1) * frame_bytes if n_up * frame_bytes > h2d else frame_bytes

            def copies():
                with torch.cuda.stream(s_up):
                    for k in range(n_up):
                        src = pinned[k % n_host]
                        if k == n_up - 1 and rem < frame_bytes:
                            rows = max(1, rem // (plan.fw * 3))
                            scratch[:rows].copy_(src[:rows], non_blocking=True)
                        else:
                            scratch.copy_(src, non_blocking=True)
                with torch.cuda.stream(s_dn):
                    out_pin.copy_(dev_out, non_blocking=True)
            copies()
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                copies()
                torch.cuda.synchronize()
            barrier()
            floor_ms = allmax(time.perf_counter() - t0) / 2 * 1e3
            del scratch, dev_out
        except Exception as e:
            floor_ms = None
            note(f"pcie floor failed: {e}")

        affinities = [None] * world
        if world > 1:
            dist.all_gather_object(affinities, numa)
        else:
            affinities = [numa]
        launches_all = int(allsum([launches])[0])

    if rank == 0:
        # dominant kernel = the launch name/level with the largest mean duration
        agg = {}
        for k in kt:
            key = (k["name"], k["level"])
            a = agg.setdefault(key, {"ms": [], "ab": k["algorithmic_bytes"]})
            if k["ms"] >= 0:
                a["ms"].append(k["ms"])
        rows = sorted(((key, float(np.mean(v["ms"])), v["ab"]) for key, v in agg.items() if v["ms"]), key=lambda r: -r[1])
        peak, peak_src = hbm_peak()
        roof = None
        if rows:
            (name, level), ms, ab = rows[0]
            ach = ab / (ms / 1e3) / 1e9
            # (the capture is of the single-GPU launch: a row band's launch moves its share of it, not measured separately)
            traffic, traffic_src = dominant_traffic(args.workload, name, level) if world == 1 else (None, "captured at N = 1 only (profiles/dominant_kernel_traffic.json)")
            step_ms = ms_total / args.steps
            roof = {"bound": "hbm", "kernel": f"{name}[level {level}]", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": ab,
                    "note": "rank 0's launches; achieved = SURVEY 8(d) algorithmic bytes of the launch / its CUDA-event time. The gather "
                            "formulation moves fewer real bytes (traffic) than the model: the kernel is issue-bound (profiles/README.md)",
                    "ms_per_launch": ms,
                    "whole_step": {"algorithmic_bytes_rank0": ab_total, "achieved": ab_total / (step_ms / 1e3) / 1e9,
                                   "frac": ab_total / (step_ms / 1e3) / 1e9 / peak, "frac_of_8TBps_nominal": ab_total / (step_ms / 1e3) / 1e9 / 8000.0},
                    "kernels": [{"kernel": f"{k[0]}[{k[1]}]", "ms": m_, "algorithmic_bytes": a_} for k, m_, a_ in rows]}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ref = CpuReference(plan, blend, bands, grid_nx(args.workload), target_seconds=20.0)
            mp, dt_ = ref.step()
            cpu = {"value": mp / dt_, "unit": UNIT, "cores": ref.cores, "kind": ref.kind, "sample": ref.sample()}
        cfg = job_config(desc, plan, roi, blend, bands, world)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": DTYPE, "data": "synthetic", "config": cfg,
                "run": {"halo": halo, "band_edges": [int(e) for e in edges], "frames_rank0": len(mine), "device_GB_rank0": info.device_bytes / 1e9,
                        "host_frames": host_mode, "host_affinity": affinities, "ms_per_composite_by_rank": ms_rank,
                        "frames_from": "procedural orthophoto evaluated per frame (synth.procedural_frame)" if procedural else "stored procedural orthophoto (synth.orthophoto)"},
                "clocks": clk, "parity": parity,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all), "steps": e2e_steps,
                        "ms_per_step": e2e_ms, "pcie_floor_ms": floor_ms,
                        "frac_of_floor": (floor_ms / e2e_ms) if floor_ms else None,
                        "floor_how": "the step's host->device and device->host byte counts copied from / to the same pinned buffers on two streams by all ranks at once, no kernels"},
                "gpu_launches": launches_all, "roofline": roof, "cpu_baseline": cpu}
        if also:
            line["also"] = also
        print(json.dumps(line), flush=True)
    cv.close()
    ok = (not parity.get("checked")) or parity.get("identical")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg2", "cfg1", "cfg3", "cfg5", "small"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--host-pool", type=int, default=0, help="pinned host frame buffers for the e2e step (0: one per frame if memory allows)")
    ap.add_argument("--bands", type=int, default=None, help="override the workload's multi-band depth (experiments)")
    ap.add_argument("--p2p", default="auto", choices=["auto", "on", "off"],
                    help="row bands exchange their level-1 halo rows over NVLink (auto / on), or recompute them (off)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
