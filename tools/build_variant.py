"""Builds a differently configured variant of the library next to the product build, for A/B runs on one GPU visit:
python tools/build_variant.py <name> -DDS_ACC_TW=64 -DDS_ACC_TH=8 ...   ->  lib/libdronestitch_cuda_<name>.so, used when DS_LIB_VARIANT=<name>."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from drone_image_stitch_cpp_b200 import build
name, extra = sys.argv[1], sys.argv[2:]
out = build.OUT.replace(".so", "_%s.so" % name)
subprocess.check_call([build.nvcc_path()] + build.NVCC_FLAGS + extra + ["-o", out, build.SRC])
print(out)
