"""Randomised parity cases: random survey geometry, transform kind, blend, bands, row-band split."""
import numpy as np

import parity_cases as P
from drone_image_stitch_cpp_b200 import synth


def random_case(lib, seed):
    rng = np.random.default_rng(seed)
    nx, ny = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    fw, fh = int(rng.integers(24, 420)), int(rng.integers(24, 320))
    ov = float(rng.uniform(0.2, 0.85))
    rot = float(rng.choice([0.0, 1.0, 3.0, 12.0, 40.0]))
    sj = float(rng.choice([0.0, 0.02, 0.2]))
    ws = float(rng.choice([1.0, 0.45, 0.37, 2.0]))
    blend = "multiband" if rng.random() < 0.75 else "feather"
    bands = int(rng.integers(0, 8))
    split = int(rng.choice([0, 2, 3]))
    kind = str(rng.choice(["plane", "plane", "plane", "affine", "homography"]))
    desc = f"seed={seed} {kind} {nx}x{ny} {fw}x{fh} ov={ov:.2f} rot={rot} sj={sj} ws={ws} {blend} bands={bands} split={split}"
    if kind == "plane":
        sv = synth.grid_survey(nx, ny, fw, fh, overlap=ov, seed=seed, rot_deg=rot, scale_jit=sj, work_scale=ws,
                               trans_jit=float(rng.uniform(0, 25)))
        specs = P.plane_specs(sv)
        if rng.random() < 0.3:
            for s in specs:
                s["affine"] = bool(rng.random() < 0.5)
    else:
        specs = P.affine_specs(seed, n=int(rng.integers(1, 5)), fw=max(fw, 60), fh=max(fh, 50), homography=(kind == "homography"))
    P.run_case(lib, specs, blend, bands, check_taps=bool(rng.random() < 0.5), band_split=split or None,
               out_format="bgra" if rng.random() < 0.2 else "bgr")
    return desc
