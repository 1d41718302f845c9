"""In-tree build of libdronestitch_cuda.so (nvcc, sm_100a only). Cross-compiles without a GPU."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "ds_runtime.cu")
DEPS = [os.path.join(_HERE, "csrc", n) for n in ("ds_runtime.cu", "ds_kernels.h", "ds_mask_kernels.h", "ds_types.h", "ds_device.h", "ds_geometry.h")]
DEPS.append(os.path.join(_HERE, "..", "include", "dronestitch.h"))
OUT = os.path.join(_HERE, "lib", "libdronestitch_cuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # every float op of the reproduced arithmetic rounds separately (no FMA contraction)
    "--shared", "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def source_hash():
    """sha256 (16 hex digits) over the kernel / runtime sources and the build flags: what a built library, a SASS listing
    or an ncu capture was made from."""
    import hashlib
    h = hashlib.sha256()
    for d in sorted(DEPS):
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; libdronestitch_cuda cannot be built")
    return p


def build_cuda(force=False, verbose=False):
    newest = max(os.path.getmtime(d) for d in DEPS)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    subprocess.check_call(cmd)
    # which sources the shipped binary was built from (the .so is git-ignored but travels to the GPU box)
    with open(OUT + ".sources", "w") as f:
        f.write(source_hash() + "\n")
    return OUT


if __name__ == "__main__":
    print(build_cuda(force=True, verbose=True))
