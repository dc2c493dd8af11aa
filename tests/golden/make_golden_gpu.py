"""Runs the REFERENCE's own CUDA kernels (oracle/_ref/libref_kernels.so, compiled from
/root/reference/models/csrc by oracle/build.py) on seeded inputs on a GPU and writes their
outputs to gpurun_out/golden/l0_reference_cuda.npz; the file is then committed under
tests/golden/ and pins oracle/kernels_oracle.c in the CPU suite (test_oracle_vs_reference_cuda).

    gpurun -- python tests/golden/make_golden_gpu.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests import _util  # noqa: E402

CASES = {
    # name: (kind, factory args)
    "fps_rand_4x4096_s1024": ("fps", lambda: _util.rand_cloud(64, 4096, 3, seed=0)[:4], 1024),
    "fps_ties_2x8192_s2048": ("fps", lambda: _util.tied_cloud(2, 8192, 3, levels=12, seed=6), 2048),
    "fps_ties_2x600_s300": ("fps", lambda: _util.tied_cloud(2, 600, 3, levels=4, seed=7), 300),
    "fps_identical_3000_s16": ("fps", lambda: torch.zeros(1, 3000, 3), 16),
}
KNN = {
    "knn_rand_1x1024x4096_k16": (3, 1024, 4096, 16, _util.rand_cloud),
    "knn_ties_1x1024x2048_k16": (3, 1024, 2048, 16, _util.tied_cloud),
    "knn_ties_1x512x1500_k32": (3, 512, 1500, 32, _util.tied_cloud),
    "knn_ties_1x200x900_k64": (3, 200, 900, 64, _util.tied_cloud),
    "knn_ties2d_1x3000x2048_k1": (2, 3000, 2048, 1, _util.tied_cloud),
    "knn_under_1x10x7_k16": (3, 10, 7, 16, _util.rand_cloud),
}


def main():
    dev = torch.device("cuda:0")
    out = {}
    for name, (_, make, S) in CASES.items():
        out[name] = _util.ref_fps(make().contiguous().to(dev), S).cpu().numpy().astype(np.int16)
    for name, (D, n, m, k, make) in KNN.items():
        inp, qry = make(1, m, D, seed=20), make(1, n, D, seed=21)
        out[name] = _util.ref_knn(inp.to(dev), qry.to(dev), k).cpu().numpy().astype(np.int16)
    dst = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "l0_reference_cuda.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
