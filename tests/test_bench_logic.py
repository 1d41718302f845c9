"""Host-side logic of bench.py that needs no GPU: which windows the in-run parity check compares (sized for the pyramid
depth, aligned, inside the canvas, straddling the band edges) and the config object both arms print."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _check_window(win, roi, bands, PH):
    x, y, w, h = win
    m = 1 << bands
    g = 8 << bands
    assert x % m == 0 and y % m == 0 and w % m == 0 and h % m == 0, (win, m)
    assert x >= 0 and y >= 0 and x + w <= roi[2] + m and y + h <= PH
    assert w > 2 * g and h > 2 * g, "nothing would be left inside the oracle's margin"


def test_parity_windows_cfg4_layout():
    roi, PH, bands = (0, 0, 86226, 31346), 31360, 5
    edges = [0, 4416, 8192, 11904, 15648, 19392, 23104, 26880, 31360]
    seen_edges = set()
    for rank in range(8):
        wins, g = bench.parity_windows(edges, rank, 8, roi, bands, PH)
        assert g == 256 and len(wins) == (1 if rank in (0, 7) else 2)
        for win, side in wins:
            _check_window(win, roi, bands, PH)
            e = edges[rank] if side == "below" else edges[rank + 1]
            assert win[1] + g < e < win[1] + win[3] - g      # the edge is inside the compared part of the window
            seen_edges.add((e, side))
    # every inner edge is checked from both sides
    assert seen_edges == {(e, s) for e in edges[1:-1] for s in ("below", "above")}
    wins, _ = bench.parity_windows([0, PH], 0, 1, roi, bands, PH)
    assert len(wins) == 1 and wins[0][1] is None
    _check_window(wins[0][0], roi, bands, PH)


def test_parity_windows_deep_pyramid():
    roi, PH, bands = (0, 0, 135466, 63573), 63744, 8
    edges = [0, 8448, 16128, 23808, 31744, 39680, 47360, 55040, 63744]
    checked = []
    for rank in range(8):
        wins, g = bench.parity_windows(edges, rank, 8, roi, bands, PH)
        assert g == 2048
        for win, side in wins:
            _check_window(win, roi, bands, PH)
            assert side == "below" and win[1] + g < edges[rank] + 256 <= win[1] + win[3] - g
            checked.append(rank)
    assert checked == [1, 4]       # two ranks, one 5120 x 4608 window each


def test_job_config_is_a_function_of_workload_and_n():
    class Plan:
        fw, fh, A = 5472, 3648, [None] * 600
    a = bench.job_config("cfg4: x", Plan, (0, 0, 86226, 31346), "multiband", 5, 8)
    b = bench.job_config("cfg4: x", Plan, (0, 0, 86226, 31346), "multiband", 5, 8)
    assert a == b and a["workload"].startswith("cfg4") and a["parallelism"].endswith("x8 of one canvas")
    assert "larger than L2" in a["l2"]
    class Small:
        fw, fh, A = 4000, 3000, [None] * 2
    assert "flushed" in bench.job_config("cfg1", Small, (0, 0, 5209, 3143), "feather", 0, 1)["l2"]
