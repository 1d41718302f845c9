"""NVLink P2P halo exchange between row-band handles. Here: several handles in one process (the emulator, and on
the GPU one device), connected with ds_p2p_export / ds_p2p_connect and driven in lock step with
ds_composite_stage - the K bands must reproduce the single-handle result bit for bit, while the level-0 feed of
every band covers only its own rows. The multi-process form (IPC handles, one process per GPU) is exercised by
bench.py under torchrun."""
import numpy as np
import pytest

from drone_image_stitch_cpp_b200 import _lib as L
from drone_image_stitch_cpp_b200 import compositor as CP
from drone_image_stitch_cpp_b200 import synth


def _exchange(lib, bands, nbands, ny=5, fw=200, fh=180, seed=51, steps=2, devices=(0,)):
    sv = synth.grid_survey(2, ny, fw, fh, overlap=0.45, seed=seed, work_scale=0.5)
    xfs = [CP.plane_transform(K, R, sv.scale) for K, R in zip(sv.Ks, sv.Rs)]
    rois = [CP.warp_roi(xf, fw, fh, lib) for xf in xfs]
    roi = CP.result_roi(rois)
    whole = CP.Canvas(roi, "multiband", bands, lib=lib)
    for i, (f, xf) in enumerate(zip(sv.frames, xfs)):
        whole.upload(i, f, xf)
    whole.composite()
    ref, refmask = whole.download()
    info = whole.info()
    m, H = 1 << info.num_bands, info.padded_height
    edges = sorted(set([0] + [((H * k // nbands) // m) * m for k in range(1, nbands)] + [H]))
    edges = [e for e in edges if e < roi[3]] + [H]
    handles = []
    for y0, y1 in zip(edges[:-1], edges[1:]):
        cb = CP.Canvas(roi, "multiband", bands, band=(y0, y1), lib=lib, device=devices[len(handles) % len(devices)])
        for i, (f, xf) in enumerate(zip(sv.frames, xfs)):
            if cb.touches(rois[i]):
                cb.upload(i, f, xf)
        handles.append(cb)
    # recomputed halos first (no exchange): launches of the level-0 feed cover band + halo
    for cb in handles:
        cb.composite()
    stacked = np.concatenate([cb.download()[0] for cb in handles], axis=0)
    assert np.array_equal(stacked, ref)
    plain_launches = [cb.info().launches_last_composite for cb in handles]
    blobs = [cb.p2p_export() for cb in handles]
    for k, cb in enumerate(handles):
        if k > 0:
            cb.p2p_connect(0, blobs[k - 1])
        if k + 1 < len(handles):
            cb.p2p_connect(1, blobs[k + 1])
    for _ in range(steps):
        for cb in handles:
            cb.composite_stage(0)
        for cb in handles:
            cb.composite_stage(1)
        for cb in handles:
            cb.synchronize()
        rows = [cb.download() for cb in handles]
        assert np.array_equal(np.concatenate([r[0] for r in rows], axis=0), ref), "exchange-mode bands differ"
        assert np.array_equal(np.concatenate([r[1] for r in rows], axis=0), refmask)
        # the exchange really ran: two hand-over launches and the pull on top of the plain schedule
        assert [cb.info().launches_last_composite for cb in handles] == [n + 3 for n in plain_launches]
    # per-frame pyramids of a straddling frame agree with the single handle where the band reads them
    # sequencing errors are reported, not silently wrong
    handles[0].composite_stage(0)
    with pytest.raises(L.DroneStitchError):
        handles[0].composite_stage(0)
    if lib.path.endswith("libdronestitch_emu.so") and len(handles) > 1:
        with pytest.raises(L.DroneStitchError):
            handles[0].composite_stage(1)      # the neighbour has not run its stage 0: the emulator cannot wait
    # leave the handles consistent: finish the step everywhere
    for cb in handles[1:]:
        cb.composite_stage(0)
    for cb in handles:
        cb.composite_stage(1)
    # back to recomputed halos
    for cb in handles:
        cb.p2p_disconnect()
    for cb in handles:
        cb.composite()
    assert np.array_equal(np.concatenate([cb.download()[0] for cb in handles], axis=0), ref)
    for cb in handles:
        cb.close()
    whole.close()


@pytest.mark.parametrize("bands,nbands", [(2, 2), (3, 3), (1, 2)])
def test_exchange_emu(emu_lib, bands, nbands):
    _exchange(emu_lib, bands, nbands)


def test_exchange_refuses_thin_bands(emu_lib):
    sv = synth.grid_survey(1, 3, 160, 160, overlap=0.4, seed=5, work_scale=0.5)
    xfs = [CP.plane_transform(K, R, sv.scale) for K, R in zip(sv.Ks, sv.Rs)]
    rois = [CP.warp_roi(xf, 160, 160, emu_lib) for xf in xfs]
    roi = CP.result_roi(rois)
    a = CP.Canvas(roi, "multiband", 4, band=(0, 16), lib=emu_lib)
    b = CP.Canvas(roi, "multiband", 4, band=(16, 32), lib=emu_lib)
    c = CP.Canvas(roi, "multiband", 4, band=(32, 10 ** 6), lib=emu_lib)
    for h in (a, b, c):
        for i, (f, xf) in enumerate(zip(sv.frames, xfs)):
            if h.touches(rois[i]):
                h.upload(i, f, xf)
    with pytest.raises(L.DroneStitchError) as e:
        c.p2p_connect(0, b.p2p_export())      # c's halo reaches beyond b into a
    assert e.value.code == L.DS_ERR_P2P_UNAVAILABLE
    with pytest.raises(L.DroneStitchError):
        a.p2p_connect(1, c.p2p_export())      # not adjacent
    for h in (a, b, c):
        h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("bands,nbands", [(3, 3), (5, 2)])
def test_exchange_gpu_one_device(cuda_lib, bands, nbands):
    # one process, one device, several handles: stream-ordered counters, plain pointers
    _exchange(cuda_lib, bands, nbands, ny=6, fw=400, fh=360)


@pytest.mark.gpu
@pytest.mark.parametrize("bands,nbands", [(3, 4), (5, 2)])
def test_exchange_gpu_two_devices(cuda_lib, bands, nbands):
    """Neighbouring bands on DIFFERENT GPUs of one process: the halo rows cross NVLink (peer access enabled by
    ds_p2p_connect), the hand-over counters live in the other device's memory. Skips on a single-GPU box; the
    multi-process form (cudaIpc handles, one process per GPU) is checked inside every bench.py --gpus N run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _exchange(cuda_lib, bands, nbands, ny=6, fw=400, fh=360, devices=(0, 1))
