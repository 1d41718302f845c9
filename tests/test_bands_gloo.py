"""N > 1 path on the CPU: two processes (gloo, world_size 2), one row band each — exactly what bench.py does
under torchrun with one process per GPU — using the tests-only emulator build of the kernels. The stacked
bands must equal the single-handle result and the oracle; no data-path collective is involved (the
all_gather below only collects the results for checking)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, emu_path, blend, bands):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from drone_image_stitch_cpp_b200 import _lib, compositor as CP, synth
    from oracle import ds_oracle as O
    lib = _lib.Library(emu_path)
    plan = synth.plan_grid(2, 5, 260, 200, overlap=0.6, seed=55, work_scale=0.5, side_overlap=0.15)
    frames = synth.cut(plan)
    xfs = [CP.plane_transform(K, R, plan.scale) for K, R in zip(plan.Ks, plan.Rs)]
    rois = [CP.warp_roi(xf, plan.fw, plan.fh, lib) for xf in xfs]
    roi = CP.result_roi(rois)
    probe = CP.Canvas(roi, blend, bands, lib=lib)
    info = probe.info()
    probe.close()
    PH, m = info.padded_height, 1 << info.num_bands
    edges = [0] + [((PH * k // world) // m) * m for k in range(1, world)] + [PH]
    cv = CP.Canvas(roi, blend, bands, band=(edges[rank], edges[rank + 1]), lib=lib)
    mine = [i for i in range(len(xfs)) if cv.touches(rois[i])]
    for i in mine:
        cv.upload(i, frames[i], xfs[i])
    cv.composite()
    pano, mask = cv.download()
    cv.close()
    parts = [None] * world
    dist.all_gather_object(parts, (pano, mask, len(mine)))
    if rank == 0:
        full = np.concatenate([p[0] for p in parts], axis=0)
        fmask = np.concatenate([p[1] for p in parts], axis=0)
        ref, refmask, roi2 = O.compose_port(frames, plan.Ks, plan.Rs, plan.scale, blend, bands)
        assert roi2 == roi
        assert np.array_equal(fmask, refmask)
        assert np.array_equal(full, ref)
        # each band received only the frames that touch it
        assert all(p[2] <= len(xfs) for p in parts) and min(p[2] for p in parts) < len(xfs)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("blend,bands", [("multiband", 3), ("feather", 0)])
def test_two_bands_two_processes(emu_lib, blend, bands):
    mp.spawn(_worker, args=(2, _free_port(), emu_lib.path, blend, bands), nprocs=2, join=True)
