"""Build recipes for the test oracles (TEST INFRASTRUCTURE ONLY).

* ``build_kernels_oracle()``: gcc-compiles ``oracle/kernels_oracle.c`` (our plain-C
  restatement of the reference's three CUDA extensions) into
  ``oracle/_build/liboracle_kernels.so``.
* ``build_reference_kernels()``: when ``/root/reference`` is present, nvcc-compiles the
  reference's OWN unmodified ``*_kernel.cu`` files (from where they lie; nothing is
  copied) together with ``oracle/ref_shim.cu`` into ``oracle/_ref/libref_kernels.so``
  for sm_100.  That library travels to the GPU box with the snapshot and is the
  bit-exact GPU oracle for FPS / k-NN indices and the "existing kernel" time bar.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_CSRC = "/root/reference/models/csrc"
ORACLE_SO = os.path.join(HERE, "_build", "liboracle_kernels.so")
REF_SO = os.path.join(HERE, "_ref", "libref_kernels.so")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_kernels_oracle(force=False):
    src = os.path.join(HERE, "kernels_oracle.c")
    if not force and not _stale(ORACLE_SO, [src]):
        return ORACLE_SO
    os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-mfma", "-fopenmp", "-shared", "-fPIC",
           "-o", ORACLE_SO, src, "-lm"]
    subprocess.check_call(cmd)
    return ORACLE_SO


def build_reference_kernels(force=False):
    """Returns the path of the built library, or None when it cannot be built here."""
    if not os.path.isdir(REFERENCE_CSRC) or shutil.which("nvcc") is None:
        return REF_SO if os.path.exists(REF_SO) else None
    srcs = [
        os.path.join(REFERENCE_CSRC, "furthest_point_sampling", "furthest_point_sampling_kernel.cu"),
        os.path.join(REFERENCE_CSRC, "k_nearest_neighbor", "k_nearest_neighbor_kernel.cu"),
        os.path.join(REFERENCE_CSRC, "correlation", "correlation_forward_kernel.cu"),
        os.path.join(REFERENCE_CSRC, "correlation", "correlation_backward_kernel.cu"),
        os.path.join(HERE, "ref_shim.cu"),
    ]
    if not force and not _stale(REF_SO, srcs):
        return REF_SO
    os.makedirs(os.path.dirname(REF_SO), exist_ok=True)
    # Same flags torch's BuildExtension would hand nvcc for TORCH_CUDA_ARCH_LIST=10.0
    # (the reference's setup.py sets none of its own): -O3, sm_100 SASS + PTX.
    cmd = ["nvcc", "-O3", "-gencode", "arch=compute_100,code=sm_100",
           "-gencode", "arch=compute_100,code=compute_100",
           "--shared", "-Xcompiler", "-fPIC", "-o", REF_SO] + srcs
    subprocess.check_call(cmd)
    return REF_SO


if __name__ == "__main__":
    print(build_kernels_oracle(force=True))
    print(build_reference_kernels(force=True))
