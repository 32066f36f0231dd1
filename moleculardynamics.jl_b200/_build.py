"""Build libmdb200.so (hand-written sm_100a CUDA + the C ABI of include/mdb200.h) in-tree with nvcc."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# MDB200_LIB / MDB200_NVCC_EXTRA: build and load tuning variants side by side (tools/tune_force.py)
LIB = os.environ.get("MDB200_LIB") or os.path.join(CSRC, "libmdb200.so")
SOURCES = ["engine.cu"]
HEADERS = ["kernels.cuh", "potentials.cuh", "rng.cuh", "slab.cuh", "small.cuh", "setup_io.cuh", "engine_setup_io.inl",
           os.path.join("..", "..", "include", "mdb200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",  # products and sums round separately, like the Julia source (see DESIGN.md "Arithmetic")
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmdb200.so cannot be built (there is no CPU fallback)")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    extra = os.environ.get("MDB200_NVCC_EXTRA", "").split()
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
