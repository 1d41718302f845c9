#!/usr/bin/env python
"""bench.py — output canvas megapixels/sec of the compositing hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (OpenCV), rank 0

Workload at N = 1: BASELINE configs[1] — a 3x3 grid of synthetic 5472x3648 drone frames, 70 % overlap,
multi-band blend with 5 bands. At N > 1 the survey grows to 3 x 3N frames (canvas N times taller) and
the canvas is cut into N row bands, one per GPU / process (weak scaling, no data-path collective).

A "step" is one ds_composite over the frames resident in HBM (value), or upload of every frame from
pinned host memory + composite + download of the whole band (e2e).
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


def workload_plan(name, n_gpus):
    from drone_image_stitch_cpp_b200 import synth
    if name == "cfg2":
        return synth.plan_grid(3, 3, 5472, 3648, overlap=0.7, seed=synth.MASTER_SEED, blocks=n_gpus), "multiband", 5, \
            (f"cfg2: {n_gpus} flight block(s) of 3x3 frames 5472x3648, 70% overlap, multi-band 5"
             + ("" if n_gpus == 1 else ", blocks stacked in y with 2% overlap, one block per GPU row band"))
    if name == "cfg1":
        return synth.plan_grid(2, 1, 4000, 3000, overlap=0.7, seed=synth.MASTER_SEED, rot_deg=1.5, blocks=n_gpus), "feather", 0, \
            f"cfg1: {n_gpus} block(s) of 2 frames 4000x3000, feather 0.02"
    if name == "cfg3":
        return synth.plan_grid(40, 3, 5472, 3648, overlap=0.7, side_overlap=0.32, seed=synth.MASTER_SEED, blocks=n_gpus), "multiband", 5, \
            f"cfg3: {n_gpus} block(s) of 3 serpentine lines x 40 frames 5472x3648, 70% forward / 32% side overlap, multi-band 5"
    if name == "small":
        return synth.plan_grid(3, 3, 912, 608, overlap=0.7, seed=synth.MASTER_SEED, blocks=n_gpus), "multiband", 5, \
            f"small: {n_gpus} block(s) of 3x3 frames 912x608, 70% overlap, multi-band 5"
    raise SystemExit(f"unknown workload {name}")


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


class CpuReference:
    """The reference's CPU path (OpenCV via oracle/cv_reference.py, else the C port) on a bounded
    sample of the workload: the same grid layout / overlap / blend with the frames scaled by an integer
    divisor so that one compose fits the time budget. Synthetic inputs are generated once (on the GPU
    when one is present — generation is not part of any timed region)."""

    def __init__(self, plan, blend, bands, budget_s, threads=None):
        from oracle import cv_reference as CR
        from oracle import ds_oracle as O
        self.plan, self.blend, self.bands = plan, blend, bands
        self.use_cv = CR.have_cv2()
        cores = os.cpu_count() or 1
        if self.use_cv:
            self.cores = CR.set_threads(threads or cores)
        else:
            O.set_threads(threads or cores)
            self.cores = O.get_threads()
        self.kind = "reference" if self.use_cv else "port"
        self.impl = ("OpenCV (cv2 wheel): AffineWarper + MultiBandBlender/FeatherBlender" if self.use_cv
                     else "oracle/ds_oracle.c (OpenMP)")
        self._prepare(6)
        mp, dt = self.step()                 # calibration + warm-up on 1/36 of the pixels
        rate = mp / dt
        div = 6
        for d in (1, 2, 3, 4, 6):
            if (mp * 36 / (d * d)) / rate <= budget_s:
                div = d
                break
        if div != 6:
            self._prepare(div)

    def _prepare(self, div):
        from drone_image_stitch_cpp_b200 import synth
        import torch
        plan = self.plan
        fw, fh = max(64, plan.fw // div), max(64, plan.fh // div)
        nx, ny = (3, 3) if self.blend == "multiband" else (2, 1)
        blocks = max(1, len(plan.A) // (nx * ny))
        self.p = synth.plan_grid(nx, ny, fw, fh, overlap=0.7, seed=plan.seed, blocks=blocks,
                                 rot_deg=3.0 if self.blend == "multiband" else 1.5, trans_jit=20.0 / div)
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        self.frames = synth.cut(self.p, None, dev)
        self.div = div

    def step(self):
        from oracle import cv_reference as CR
        from oracle import ds_oracle as O
        t0 = time.perf_counter()
        if self.use_cv:
            pano, _, roi = CR.compose_cv2(self.frames, self.p.Ks, self.p.Rs, self.p.scale, self.blend, self.bands)
        else:
            pano, _, roi = O.compose_port(self.frames, self.p.Ks, self.p.Rs, self.p.scale, self.blend, self.bands)
        dt = time.perf_counter() - t0
        self.last_dt = dt
        return roi[2] * roi[3] / 1e6, dt

    def sample(self):
        return (f"{len(self.frames)} frames of {self.p.fw}x{self.p.fh} (1/{self.div} linear scale of the workload), "
                f"{self.impl}, one compose = {self.last_dt:.2f} s")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    plan, blend, bands, desc = workload_plan(args.workload, args.gpus)
    total = args.steps + args.warmup
    budget = max(1.0, 150.0 / max(total, 1))
    ref = CpuReference(plan, blend, bands, budget)
    kind, cores = ref.kind, ref.cores
    mps, secs = 0.0, 0.0
    nsteps = 0
    t_all = time.perf_counter()
    for i in range(total):
        mp, dt = ref.step()
        if i >= args.warmup:
            mps += mp; secs += dt; nsteps += 1
        if time.perf_counter() - t_all > 240 and nsteps >= 1:
            break
    samples = ref.sample()
    v = mps / secs
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": nsteps,
            "warmup": args.warmup, "ms_per_step": secs / nsteps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int16/f32", "data": "synthetic", "config": {"workload": desc, "sample": samples},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": samples},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cores local to GPU {idx}"
    except Exception as e:  # measurement nicety only
        return f"unbound ({type(e).__name__})"
    return "unbound"


def run_native(args):
    import torch
    import torch.distributed as dist
    from drone_image_stitch_cpp_b200 import _lib, compositor as CP, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (libdronestitch_cuda has no CPU fallback)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.default_library()

    plan, blend, bands, desc = workload_plan(args.workload, world)
    if args.bands is not None and blend == "multiband":
        bands = args.bands
        desc += f" [bands overridden: {bands}]"
    xfs = [CP.plane_transform(K, R, plan.scale) for K, R in zip(plan.Ks, plan.Rs)]
    rois = [CP.warp_roi(xf, plan.fw, plan.fh, lib) for xf in xfs]
    roi = CP.result_roi(rois)
    # row bands: edges at multiples of 2^bands
    probe = CP.Canvas(roi, blend, bands, lib=lib, device=local)
    pinfo = probe.info()
    probe.close()
    PH, m = pinfo.padded_height, 1 << pinfo.num_bands
    edges = [0] + [((PH * k // world) // m) * m for k in range(1, world)] + [PH]
    band = (edges[rank], edges[rank + 1]) if world > 1 else None

    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        cv = CP.Canvas(roi, blend, bands, out_format="bgr", device=local, band=band, stream=stream.cuda_stream, lib=lib)
        mine = [i for i in range(len(xfs)) if band is None or cv.touches(rois[i])]
        frames_dev = synth.cut(plan, mine, device=f"cuda:{local}", as_torch=True)
        pinned = []
        for f in frames_dev:
            h = torch.empty(f.shape, dtype=torch.uint8, pin_memory=True)
            h.copy_(f)
            pinned.append(h)
        del frames_dev
        torch.cuda.synchronize()
        host_np = [p.numpy() for p in pinned]
        for i, arr in zip(mine, host_np):
            cv.upload(i, arr, xfs[i])
        info = cv.info()
        out_rows = min(info.band_y1, roi[3]) - info.band_y0
        out_pin = torch.empty((out_rows, roi[2], 3), dtype=torch.uint8, pin_memory=True)

        # neighbouring bands hand each other the level-1 halo rows over NVLink (no collective on the data path;
        # torch.distributed only carries the 64-byte IPC handles once, here)
        # Measured on cfg2 (5 bands, halo of ~280 rows): recomputing the halo costs 0.05-0.18 ms per composite, the
        # exchange 0.06 ms of pull per edge plus two cross-GPU hand-overs - a wash. From 6 bands up (halo >= 560
        # rows, doubling per band) the exchange wins, so "auto" connects only then.
        halo = "recomputed per band (no exchange)"
        use_p2p = args.p2p == "on" or (args.p2p == "auto" and info.num_bands >= 6)
        if world > 1 and blend == "multiband" and use_p2p:
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
                halo = "NVLink P2P pull of the level-1 halo rows (ds_p2p_connect), recomputed in the pipelined e2e schedule"
            else:
                cv.p2p_disconnect()
                halo = "recomputed per band (P2P unavailable: " + next(o[1] for o in oks if not o[0])[:120] + ")"

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---------------- value: composite-only, frames resident in HBM
        for _ in range(max(args.warmup, 3)):
            cv.composite_async()
        cv.set_profiling(True)
        barrier()
        clocks = ClockSampler(local)
        clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            cv.composite_async()
        e1.record(stream)
        barrier()
        clk = clocks.stop()
        ms_total = e0.elapsed_time(e1)
        kt = cv.kernel_times()
        cv.set_profiling(False)
        launches = int(cv.info().launches_last_composite) * args.steps
        t = torch.tensor([ms_total], device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        canvas_mp = roi[2] * roi[3] / 1e6
        value = canvas_mp * args.steps / (ms_total / 1e3)

        # ---------------- e2e: upload from pinned host + composite + download, every step
        e2e_steps = max(2, min(args.steps, 5))
        # the step's result is the panorama (what composePanorama hands back, stitch_robust.cpp:256); the
        # result mask stays on the device unless asked for
        d2h = int(out_pin.numel())

        def e2e_step():
            # the caller's frames stay valid for the step, so the uploads are queued (DS_UPLOAD_ASYNC), the composite
            # works through the canvas in row slices as the frames arrive and the download copies every slice out
            # as soon as it is final: host->device copies, kernels and device->host copies overlap
            for i, arr in zip(mine, host_np):
                cv.upload(i, arr, xfs[i], async_=True)
            cv.composite_async()
            cv.download(out=out_pin.numpy(), want_mask=False)
            cv.synchronize()

        e2e_step()
        barrier()
        h2d0 = int(cv.info().h2d_bytes_total)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        # bytes the library actually copied (a row-band handle only pulls the source rows its band reads)
        h2d = (int(cv.info().h2d_bytes_total) - h2d0) // e2e_steps
        t = torch.tensor([dt], device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_val = canvas_mp * e2e_steps / float(t.item())
        h2d_t = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(h2d_t)
        ab_total = int(cv.info().algorithmic_bytes)

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
            traffic = None
            tp = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
            if os.path.exists(tp):
                try:
                    traffic = json.load(open(tp)).get(f"{args.workload}:{name}:{level}")
                except Exception:
                    traffic = None
            roof = {"bound": "hbm", "kernel": f"{name}[level {level}]", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": ab,
                    "note": "achieved = SURVEY 8(d) algorithmic bytes / CUDA-event time; the gather formulation moves fewer real bytes "
                            "(traffic) than the model, the kernel is issue-bound (profiles/README.md)",
                    "ms_per_launch": ms,
                    "whole_step": {"algorithmic_bytes": ab_total, "achieved": ab_total / (ms_total / args.steps / 1e3) / 1e9,
                                   "frac": ab_total / (ms_total / args.steps / 1e3) / 1e9 / peak},
                    "kernels": [{"kernel": f"{k[0]}[{k[1]}]", "ms": m_, "algorithmic_bytes": a_} for k, m_, a_ in rows]}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ref = CpuReference(plan, blend, bands, budget_s=25.0)
            mp, dt = ref.step()
            cpu = {"value": mp / dt, "unit": UNIT, "cores": ref.cores, "kind": ref.kind, "sample": ref.sample()}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8/int16/f32", "data": "synthetic",
                "config": {"workload": desc, "canvas": [roi[2], roi[3]], "frames": len(xfs), "bands_per_gpu": 1,
                           "parallelism": f"row-bands x{world}", "halo": halo, "host_affinity": numa, "l2": "inputs larger than L2 (720 MB of BGRX frames per band)"},
                "clocks": clk,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d_t[0].item()),
                        "d2h_bytes_per_step": int(h2d_t[1].item()), "steps": e2e_steps},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    cv.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg1", "cfg3", "small"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bands", type=int, default=None, help="override the workload's multi-band depth (experiments)")
    ap.add_argument("--p2p", default="auto", choices=["auto", "on", "off"],
                    help="row bands exchange their level-1 halo rows over NVLink (on), recompute them (off), or decide by pyramid depth (auto)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
