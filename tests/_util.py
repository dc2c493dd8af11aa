"""Shared helpers for the tests: seeded input generators and ctypes access to the
oracles.  Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may
touch oracle/ -- the product never does."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import build as oracle_build  # noqa: E402

_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        _oracle = ctypes.CDLL(oracle_build.build_kernels_oracle())
    return _oracle


def ref_lib():
    """The reference's own CUDA kernels (oracle/_ref), or None when not built."""
    global _ref
    if _ref is None:
        path = oracle_build.build_reference_kernels()
        if path is None or not os.path.exists(path):
            return None
        _ref = ctypes.CDLL(path)
    return _ref


def _np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------- C oracle (numpy in/out)
def oracle_fps(xyz, n_samples):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    out = np.empty((B, n_samples), dtype=np.int64)
    oracle_lib().oracle_furthest_point_sampling(_np_ptr(xyz), B, N, n_samples, _np_ptr(out))
    return out


def oracle_knn(input_xyz, query_xyz, k):
    input_xyz = np.ascontiguousarray(input_xyz, dtype=np.float32)
    query_xyz = np.ascontiguousarray(query_xyz, dtype=np.float32)
    B, n, D = query_xyz.shape
    m = input_xyz.shape[1]
    out = np.empty((B, n, k), dtype=np.int64)
    oracle_lib().oracle_k_nearest_neighbor(B, n, m, k, D, _np_ptr(query_xyz), _np_ptr(input_xyz), _np_ptr(out))
    return out


def oracle_corr_fwd(in1_nhwc, in2_nhwc, md):
    in1 = np.ascontiguousarray(in1_nhwc, dtype=np.float32)
    in2 = np.ascontiguousarray(in2_nhwc, dtype=np.float32)
    B, H, W, C = in1.shape
    out = np.empty((B, (2 * md + 1) ** 2, H, W), dtype=np.float32)
    oracle_lib().oracle_correlation_forward(_np_ptr(out), _np_ptr(in1), _np_ptr(in2), B, C, H, W, md)
    return out


def oracle_corr_bwd(gout, in1_nhwc, in2_nhwc, md):
    gout = np.ascontiguousarray(gout, dtype=np.float32)
    in1 = np.ascontiguousarray(in1_nhwc, dtype=np.float32)
    in2 = np.ascontiguousarray(in2_nhwc, dtype=np.float32)
    B, H, W, C = in1.shape
    g1 = np.empty((B, C, H, W), dtype=np.float32)
    g2 = np.empty((B, C, H, W), dtype=np.float32)
    oracle_lib().oracle_correlation_backward(_np_ptr(gout), _np_ptr(g1), _np_ptr(g2), _np_ptr(in1), _np_ptr(in2),
                                             B, C, H, W, md)
    return g1, g2


# ---------------------------------------------------------------- reference CUDA kernels (torch cuda in/out)
def _tp(t):
    return ctypes.c_void_p(t.data_ptr())


def ref_fps(xyz, n_samples):
    B, N, _ = xyz.shape
    out = torch.empty((B, n_samples), dtype=torch.int64, device=xyz.device)
    tmp = torch.ones((B, N), dtype=torch.float32, device=xyz.device) * 1e10   # furthest_point_sampling.cpp:12
    torch.cuda.synchronize()
    code = ref_lib().ref_fps(_tp(xyz), _tp(tmp), B, N, n_samples, _tp(out))
    torch.cuda.synchronize()
    assert code == 0, code
    return out


def ref_knn(input_xyz, query_xyz, k):
    B, n, D = query_xyz.shape
    m = input_xyz.shape[1]
    out = torch.zeros((B, n, k), dtype=torch.int64, device=query_xyz.device)   # k_nearest_neighbor.cpp:16
    torch.cuda.synchronize()
    code = ref_lib().ref_knn(B, n, m, k, D, _tp(query_xyz), _tp(input_xyz), _tp(out))
    torch.cuda.synchronize()
    assert code == 0, code
    return out


def ref_corr_fwd(in1, in2, md):
    B, H, W, C = in1.shape
    out = torch.zeros((B, (2 * md + 1) ** 2, H, W), dtype=torch.float32, device=in1.device)  # correlation.cpp:17
    torch.cuda.synchronize()
    code = ref_lib().ref_corr_fwd(_tp(out), _tp(in1), _tp(in2), B, C, H, W, md)
    torch.cuda.synchronize()
    assert code == 0, code
    return out


def ref_corr_bwd(gout, in1, in2, md):
    B, H, W, C = in1.shape
    g1 = torch.empty((B, C, H, W), dtype=torch.float32, device=in1.device)
    g2 = torch.empty((B, C, H, W), dtype=torch.float32, device=in1.device)
    torch.cuda.synchronize()
    code = ref_lib().ref_corr_bwd(_tp(gout), _tp(g1), _tp(g2), _tp(in1), _tp(in2), B, C, H, W, md)
    torch.cuda.synchronize()
    assert code == 0, code
    return g1, g2


# ---------------------------------------------------------------- seeded inputs
def rand_cloud(B, N, D=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand((B, N, D), generator=g)


def tied_cloud(B, N, D=3, levels=6, seed=0):
    """Coordinates on a coarse lattice (exactly representable) => masses of exact distance
    ties and duplicate points, the case the datasets produce by sampling with replacement."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, levels, (B, N, D), generator=g).float() * 0.25


def synthetic_pc(B, N, seed=0):
    """SURVEY 8(d) / BASELINE.md 3 generator: u~U[0,959], v~U[0,539], z~U[5,35], f=1050."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand((B, N), generator=g) * 959.0
    v = torch.rand((B, N), generator=g) * 539.0
    z = torch.rand((B, N), generator=g) * 30.0 + 5.0
    return torch.stack([(u - 479.5) * z / 1050.0, (v - 269.5) * z / 1050.0, z], dim=-1)
