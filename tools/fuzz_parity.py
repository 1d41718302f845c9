"""Randomised parity fuzzing against the oracle (tests/fuzz_cases.py).
usage: python tools/fuzz_parity.py [--gpu] [iterations] [seed]     (--gpu: the CUDA library instead of the emulator)"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from drone_image_stitch_cpp_b200 import _lib  # noqa: E402
from fuzz_cases import random_case  # noqa: E402

args = [a for a in sys.argv[1:] if a != "--gpu"]
if "--gpu" in sys.argv:
    lib = _lib.default_library()          # libdronestitch_cuda on a B200
else:
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu"), "-s"])
    lib = _lib.Library(os.path.join(ROOT, "tests", "emu", "_build", "libdronestitch_emu.so"))
n = int(args[0]) if len(args) > 0 else 50
seed0 = int(args[1]) if len(args) > 1 else 1000
t0 = time.time()
fails = 0
for it in range(n):
    try:
        desc = random_case(lib, seed0 + it)
    except Exception as e:  # noqa: BLE001
        fails += 1
        print("FAIL seed", seed0 + it, "->", type(e).__name__, str(e)[:300], flush=True)
        continue
    if it % 10 == 0:
        print(f"ok {it} ({time.time() - t0:.0f}s) {desc}", flush=True)
print("done", n, "iterations,", fails, "failures, in", round(time.time() - t0), "s")
sys.exit(1 if fails else 0)
