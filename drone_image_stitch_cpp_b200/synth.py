"""Synthetic survey generator: a procedural orthophoto and drone frames cut from
it through known affine transforms, so overlaps are photometrically consistent.

Used by tests/ and bench.py only to make inputs; torch is plumbing here (runs on CPU or on the GPU).
Cameras follow what cv::Stitcher holds in SCANS mode (reference: src/stitch_app.cpp:208, :259):
K = diag(a, a, 1) float32, R = 3x3 float32 affine, warp scale = a (a = 1/work_scale).
"""
import math
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch
import torch.nn.functional as F

MASTER_SEED = 20261018


def orthophoto(h, w, seed=MASTER_SEED, device="cpu"):
    """fBm value noise + a few flat shapes, uint8 BGR (h, w, 3), mean ~110, sigma ~50."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    img = torch.zeros(1, 3, h, w, device=device)
    amp, tot = 1.0, 0.0
    for octave in range(7):
        gh = max(2, math.ceil(h / (512 >> octave)) + 2)
        gw = max(2, math.ceil(w / (512 >> octave)) + 2)
        n = torch.rand(1, 3, gh, gw, generator=g).to(device)
        img += amp * F.interpolate(n, size=(h, w), mode="bicubic", align_corners=True)
        tot += amp
        amp *= 0.6
    img = (img / tot - 0.5) * 2.0 + 0.43
    # flat shapes (fields / roofs): axis-aligned rectangles with random tint
    nshape = max(4, (h * w) // 400_000)
    ys = torch.randint(0, h, (nshape,), generator=g).tolist()
    xs = torch.randint(0, w, (nshape,), generator=g).tolist()
    hs = torch.randint(20, max(21, min(h, 600)), (nshape,), generator=g).tolist()
    ws = torch.randint(20, max(21, min(w, 600)), (nshape,), generator=g).tolist()
    tint = (torch.rand(nshape, 3, generator=g) * 0.5 - 0.25).to(device)
    for i in range(nshape):
        img[0, :, ys[i]:ys[i] + hs[i], xs[i]:xs[i] + ws[i]] += tint[i].view(3, 1, 1)
    fine = torch.rand(1, 1, h, w, generator=g).to(device) * 0.08 - 0.04
    img = (img + fine).clamp_(0.0, 1.0)
    return (img[0].permute(1, 2, 0) * 255.0).round_().to(torch.uint8).contiguous()


@dataclass
class Survey:
    frames: List[np.ndarray]            # HxWx3 uint8 BGR
    Ks: List[np.ndarray]                # 3x3 float32
    Rs: List[np.ndarray]                # 3x3 float32 affine
    scale: float
    A: List[np.ndarray] = field(default_factory=list)   # 2x3 float64 frame px -> ortho px (ground truth)


def cut_frame(ortho_chw, A, fw, fh):
    """Sample the orthophoto (1,3,H,W float) at A*(x,y,1) for every frame pixel, bilinear."""
    H, W = ortho_chw.shape[-2:]
    dev = ortho_chw.device
    ys, xs = torch.meshgrid(torch.arange(fh, device=dev, dtype=torch.float32),
                            torch.arange(fw, device=dev, dtype=torch.float32), indexing="ij")
    ox = A[0, 0] * xs + A[0, 1] * ys + A[0, 2]
    oy = A[1, 0] * xs + A[1, 1] * ys + A[1, 2]
    grid = torch.stack([(ox + 0.5) / W * 2 - 1, (oy + 0.5) / H * 2 - 1], dim=-1)[None]
    out = F.grid_sample(ortho_chw, grid, mode="bilinear", padding_mode="reflection", align_corners=False)
    return out[0].permute(1, 2, 0).round_().clamp_(0, 255).to(torch.uint8).contiguous()


@dataclass
class SurveyPlan:
    """Transforms of a grid survey without the pixels (cheap); `cut` makes the frames."""
    fw: int
    fh: int
    ortho_w: int
    ortho_h: int
    seed: int
    Ks: List[np.ndarray]
    Rs: List[np.ndarray]
    A: List[np.ndarray]
    scale: float


def plan_grid(nx, ny, fw, fh, overlap=0.7, seed=MASTER_SEED, rot_deg=3.0, scale_jit=0.02, trans_jit=20.0,
              work_scale=1.0, serpentine=True, side_overlap=None, blocks=1, block_overlap=0.02):
    """nx x ny frames of fw x fh, step = (1-overlap) of the frame size, with jitter.
    work_scale != 1 exercises the K = diag(1/ws), scale = 1/ws camera convention."""
    rng = np.random.default_rng(seed)
    so = overlap if side_overlap is None else side_overlap
    stepx, stepy = fw * (1.0 - overlap), fh * (1.0 - so)
    margin = int(0.15 * max(fw, fh)) + 64
    # `blocks` flight blocks of nx x ny frames stacked in y, adjacent blocks overlapping by
    # block_overlap * fh (weak-scaling workload of bench.py: one block per GPU row band)
    block_h = stepy * (ny - 1) + fh
    block_step = block_h - block_overlap * fh
    W = int(stepx * (nx - 1) + fw) + 2 * margin
    H = int(block_step * (blocks - 1) + block_h) + 2 * margin
    a = np.float32(1.0 / work_scale)
    Ks, Rs, As = [], [], []
    for jj in range(ny * blocks):
        blk, j = divmod(jj, ny)
        cols = range(nx) if (not serpentine or jj % 2 == 0) else range(nx - 1, -1, -1)
        for i in cols:
            th = math.radians(rng.uniform(-rot_deg, rot_deg))
            s = rng.uniform(1 - scale_jit, 1 + scale_jit)
            tx = margin + i * stepx + rng.uniform(-trans_jit, trans_jit)
            ty = margin + blk * block_step + j * stepy + rng.uniform(-trans_jit, trans_jit)
            cx, cy = fw / 2.0, fh / 2.0
            # rotate/scale about the frame centre, then translate
            r00, r01, r10, r11 = s * math.cos(th), -s * math.sin(th), s * math.sin(th), s * math.cos(th)
            A = np.array([[r00, r01, tx + cx - (r00 * cx + r01 * cy)],
                          [r10, r11, ty + cy - (r10 * cx + r11 * cy)]], np.float64)
            As.append(A)
            Ks.append(np.array([[a, 0, 0], [0, a, 0], [0, 0, 1]], np.float32))
            # AffineWarper's backward map is x_src = a * L * (u/a + L^T T0) with R = [L | T0] (OpenCV
            # PlaneProjector::mapBackward after getRTfromHomogeneous), i.e. R holds the canvas->frame
            # map in work-scale units. Solve L, T0 so that it equals the inverse of A.
            Bl = np.linalg.inv(A[:, :2])
            Bt = -Bl @ A[:, 2]
            T0 = np.linalg.solve(Bl @ Bl.T, Bt) / float(a)
            Rs.append(np.array([[Bl[0, 0], Bl[0, 1], T0[0]], [Bl[1, 0], Bl[1, 1], T0[1]], [0, 0, 1]], np.float32))
    return SurveyPlan(fw, fh, W, H, seed, Ks, Rs, As, float(a))


def cut(plan, indices=None, device="cpu", as_torch=False, ortho=None):
    """Cut the frames `indices` (default all) of a plan out of its orthophoto. Returns a list aligned
    with `indices` of HxWx3 uint8 arrays (numpy, or torch tensors on `device` when as_torch)."""
    if ortho is None:
        ortho = orthophoto(plan.ortho_h, plan.ortho_w, plan.seed, device)
    ortho_f = ortho.permute(2, 0, 1)[None].float()
    idx = range(len(plan.A)) if indices is None else indices
    out = []
    for i in idx:
        fr = cut_frame(ortho_f, torch.tensor(plan.A[i], dtype=torch.float32), plan.fw, plan.fh)
        out.append(fr if as_torch else fr.cpu().numpy())
    return out


def grid_survey(nx, ny, fw, fh, overlap=0.7, seed=MASTER_SEED, rot_deg=3.0, scale_jit=0.02, trans_jit=20.0,
                work_scale=1.0, device="cpu", serpentine=True, side_overlap=None):
    plan = plan_grid(nx, ny, fw, fh, overlap, seed, rot_deg, scale_jit, trans_jit, work_scale, serpentine, side_overlap)
    return Survey(cut(plan, None, device), plan.Ks, plan.Rs, plan.scale, plan.A)


# ---------------------------------------------------------------------------------------------
# Procedural orthophoto for surveys whose canvas is too large to store (BASELINE configs 3-5): the ground colour is a
# closed-form function of the orthophoto coordinate - a sum of plane waves over eight octaves (fields, roads, texture
# down to a 3 px period) plus a few hundred soft-edged "parcels" from a hashed lattice - so a frame is evaluated directly
# at the ortho position of each of its pixels. Same statistics as `orthophoto` (mean ~110, sigma ~45, fine texture of
# +-10 levels so Laplacian levels are never trivially zero); overlaps are photometrically consistent by construction.

def _wave_bank(seed):
    rng = np.random.default_rng(int(seed) + 77)
    waves = []
    for octave in range(8):
        period = 1800.0 / (2.3 ** octave)              # 1800 px ... 5.3 px
        amp = 26.0 * (0.72 ** octave)
        for _ in range(3):
            th = rng.uniform(0, 2 * math.pi)
            k = 2 * math.pi / (period * rng.uniform(0.8, 1.25))
            waves.append((k * math.cos(th), k * math.sin(th), rng.uniform(0, 2 * math.pi, 3), amp * rng.uniform(0.6, 1.0, 3)))
    for _ in range(4):                                  # the finest texture: 3-4 px periods, +-4 levels each
        th = rng.uniform(0, 2 * math.pi)
        k = 2 * math.pi / rng.uniform(3.0, 4.2)
        waves.append((k * math.cos(th), k * math.sin(th), rng.uniform(0, 2 * math.pi, 3), np.full(3, 4.0)))
    return waves


_WAVES = {}


def procedural_frame(plan, i, device="cpu", as_torch=False):
    """Frame i of `plan` (HxWx3 uint8 BGR) evaluated from the procedural orthophoto through its ground-truth map A_i."""
    waves = _WAVES.setdefault(plan.seed, _wave_bank(plan.seed))
    A = plan.A[i]
    fw, fh = plan.fw, plan.fh
    dt = torch.float32
    ys, xs = torch.meshgrid(torch.arange(fh, device=device, dtype=dt), torch.arange(fw, device=device, dtype=dt), indexing="ij")
    # ortho coordinates in float64-accurate form: large offsets are folded into the phases per wave
    ox = float(A[0, 0]) * xs + float(A[0, 1]) * ys
    oy = float(A[1, 0]) * xs + float(A[1, 1]) * ys
    tx, ty = float(A[0, 2]), float(A[1, 2])
    out = torch.full((3, fh, fw), 112.0, device=device, dtype=dt)
    for kx, ky, ph, amp in waves:
        base = kx * ox + ky * oy
        off = math.fmod(kx * tx + ky * ty, 2 * math.pi)
        for c in range(3):
            out[c] += float(amp[c]) * torch.sin(base + (off + float(ph[c])))
    # parcels: a 700 px lattice, each cell tinted by a hash of its index (soft 40 px edges)
    gx, gy = (ox + tx) / 700.0, (oy + ty) / 700.0
    cx, cy = torch.floor(gx), torch.floor(gy)
    hsh = torch.frac(torch.sin(cx * 127.1 + cy * 311.7 + (plan.seed % 1000) * 0.37) * 43758.5453)
    edge = torch.minimum(torch.minimum(gx - cx, cx + 1 - gx), torch.minimum(gy - cy, cy + 1 - gy)).clamp_(0, 0.06) / 0.06
    tint = (hsh - 0.5) * 60.0 * edge
    out[0] += tint * 0.6
    out[1] += tint
    out[2] += tint * 0.8
    img = out.clamp_(0.0, 255.0).round_().to(torch.uint8).permute(1, 2, 0).contiguous()
    return img if as_torch else img.cpu().numpy()


def plan_survey(name, seed=MASTER_SEED):
    """The BASELINE.json survey layouts: (plan, blend, bands, description)."""
    fw, fh = 5472, 3648
    if name == "cfg2":
        return plan_grid(3, 3, fw, fh, overlap=0.7, seed=seed), "multiband", 5, "cfg2: 3x3 grid of 5472x3648 frames, 70% overlap, multi-band 5"
    if name == "cfg3":
        return (plan_grid(40, 3, fw, fh, overlap=0.7, side_overlap=0.32, seed=seed), "multiband", 5,
                "cfg3: 120-frame serpentine flight, 3 lines x 40 frames of 5472x3648, 70% forward / 32% side overlap, multi-band 5")
    if name == "cfg4":
        return (plan_grid(50, 12, fw, fh, overlap=0.7, side_overlap=0.32, seed=seed), "multiband", 5,
                "cfg4: 600-frame survey, 12 lines x 50 frames of 5472x3648, 70% forward / 32% side overlap, multi-band 5")
    if name == "cfg5":
        return (plan_grid(80, 25, fw, fh, overlap=0.7, side_overlap=0.32, seed=seed), "multiband", 8,
                "cfg5: 2000-frame large-area survey, 25 lines x 80 frames of 5472x3648, 70% forward / 32% side overlap, multi-band 8")
    if name == "cfg1":
        return plan_grid(2, 1, 4000, 3000, overlap=0.7, seed=seed, rot_deg=1.5), "feather", 0, "cfg1: 2 frames of 4000x3000, feather 0.02"
    if name == "small":
        return (plan_grid(6, 4, 640, 480, overlap=0.7, side_overlap=0.32, seed=seed), "multiband", 5,
                "small: 4 lines x 6 frames of 640x480 (test-sized stand-in for cfg4), multi-band 5")
    raise ValueError(f"unknown survey {name}")
