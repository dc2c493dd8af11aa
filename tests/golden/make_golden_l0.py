"""Generates tests/golden/l0_reference_py.npz by running the REFERENCE's own Python
implementations (models/csrc/wrapper.py fallbacks, imported from /root/reference) on the
reference's self-test recipes (models/csrc/*/*_test.cpp), scaled so the fixture stays small.

Run here (needs /root/reference; not needed on the GPU box):
    python tests/golden/make_golden_l0.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import _util  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_wrapper", "/root/reference/models/csrc/wrapper.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)   # prints the reference's "Failed to load CUDA extensions" notice: expected

out = {}

# FPS: recipe of furthest_point_sampling_test.cpp:34-44 (seed 0, rand[64,4096,3], 1024 samples);
# first 16 clouds kept.
xyz = _util.rand_cloud(64, 4096, 3, seed=0)[:16]
out["fps_rand_16x4096_s1024"] = ref.furthest_point_sampling(xyz, 1024, cpp_impl=False).numpy().astype(np.int16)

# FPS on the synthetic camera-frustum generator at the model's size.
pc = _util.synthetic_pc(2, 8192, seed=0)
out["fps_synth_2x8192_s4096"] = ref.furthest_point_sampling(pc, 4096, cpp_impl=False).numpy().astype(np.int16)

# kNN: recipe of k_nearest_neighbor_test.cpp:25-38 (seed 0, rand[8,8192,3] inputs and queries, k=16);
# first 2 batches / first 2048 queries kept.
inp = _util.rand_cloud(8, 8192, 3, seed=0)[:2]
qry = _util.rand_cloud(8, 8192, 3, seed=1)[:2, :2048]
out["knn_rand_2x2048x8192_k16"] = ref.k_nearest_neighbor(inp, qry, 16, cpp_impl=False).numpy().astype(np.int16)

# 2-D kNN, k=1 (the CLFM call, clfm.py:60)
inp2 = _util.rand_cloud(1, 2048, 2, seed=2) * 100
qry2 = _util.rand_cloud(1, 4000, 2, seed=3) * 100
out["knn2d_rand_1x4000x2048_k1"] = ref.k_nearest_neighbor(inp2, qry2, 1, cpp_impl=False).numpy().astype(np.int16)

# PWC correlation: recipe of correlation_test.cpp:45-60 scaled down (seed 0, rand, d=4), fwd + autograd grads.
g = torch.Generator().manual_seed(0)
a = torch.rand((2, 32, 20, 36), generator=g, requires_grad=True)
b = torch.rand((2, 32, 20, 36), generator=g, requires_grad=True)
go = torch.rand((2, 81, 20, 36), generator=g)
o = ref.correlation2d(a, b, 4, cpp_impl=False)
o.backward(go)
out["corr_fwd"] = o.detach().numpy()
out["corr_g1"] = a.grad.numpy()
out["corr_g2"] = b.grad.numpy()

np.savez_compressed(os.path.join(HERE, "l0_reference_py.npz"), **out)
print({k: v.shape for k, v in out.items()})
