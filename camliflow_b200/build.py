"""In-tree build of libcamli_b200.so (hand-written CUDA for sm_100a, plain nvcc).

The library has no torch / pybind dependency: it is a C-ABI shared object (see
include/camli_b200.h) that the Python host side loads with ctypes.  nvcc
cross-compiles it without a GPU in a few seconds; the .so stays in-tree
(camliflow_b200/_build/) so it travels with a repo snapshot.
"""
import glob
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
# CAMLI_LIB_PATH: load / build another copy of the library (A/B experiments with -D switches, scripts/build_variants.sh)
LIB_PATH = os.environ.get("CAMLI_LIB_PATH") or os.path.join(PKG_DIR, "_build", "libcamli_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC_DIR, "*.cu")))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC_DIR, "*.cuh")) + \
        glob.glob(os.path.join(PKG_DIR, "..", "include", "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build_library(force=False, verbose=False):
    """Compile every .cu under camliflow_b200/csrc into one shared library."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libcamli_b200.so")
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("CAMLI_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB_PATH] + sources()
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
