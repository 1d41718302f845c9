"""Shared comparison helpers for the parity tests."""
import numpy as np


def lsb_stats(a, b):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return dict(max=int(d.max()) if d.size else 0, frac_le1=float((d <= 1).mean()) if d.size else 1.0,
                n_diff=int((d != 0).sum()))


def assert_blend_parity(pano, ref, exact=True):
    """north_star bar: <= 1 LSB on >= 99.99 % of pixels; this implementation is expected bit-exact."""
    st = lsb_stats(pano, ref)
    assert st["frac_le1"] >= 0.9999, st
    if exact:
        assert st["n_diff"] == 0, st
    return st
