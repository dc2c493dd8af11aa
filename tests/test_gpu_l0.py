"""GPU parity of the three native ops, called through the C-ABI (camliflow_b200.csrc ->
libcamli_b200.so), against (a) the C oracle and (b) the reference's own unmodified CUDA kernels
(oracle/_ref/libref_kernels.so, built from /root/reference by oracle/build.py).
Integer outputs must be bit-exact; the cost volume within the reference test's 1e-6 mean |diff|."""
import numpy as np
import pytest
import torch

from tests import _util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _csrc():
    from camliflow_b200 import csrc
    return csrc


def _have_ref():
    return _util.ref_lib() is not None


FPS_CASES = [
    # name, cloud factory, S
    ("selftest_recipe_8x4096", lambda: _util.rand_cloud(64, 4096, 3, seed=0)[:8], 1024),
    ("model_size_2x8192", lambda: _util.synthetic_pc(2, 8192, seed=0), 4096),
    ("batch8_8192", lambda: _util.synthetic_pc(8, 8192, seed=3), 4096),
    ("n300", lambda: _util.rand_cloud(3, 300, 3, seed=1), 100),
    ("n1025", lambda: _util.rand_cloud(2, 1025, 3, seed=2), 1024),
    ("n3000", lambda: _util.rand_cloud(2, 3000, 3, seed=4), 777),
    ("n5000", lambda: _util.rand_cloud(2, 5000, 3, seed=5), 2000),
    ("ties_lattice_8192", lambda: _util.tied_cloud(2, 8192, 3, levels=12, seed=6), 4096),
    ("ties_lattice_600", lambda: _util.tied_cloud(2, 600, 3, levels=4, seed=7), 300),
    ("all_identical", lambda: torch.zeros(1, 3000, 3), 16),
    ("duplicates_with_replacement", lambda: _util.synthetic_pc(1, 5000, seed=8)[:, torch.randint(
        0, 5000, (8192,), generator=torch.Generator().manual_seed(9))], 4096),
    ("streaming_n10000", lambda: _util.synthetic_pc(2, 10000, seed=10), 1500),
    ("n16384", lambda: _util.synthetic_pc(1, 16384, seed=12), 1000),
    ("streaming_n17000", lambda: _util.synthetic_pc(1, 17000, seed=13), 600),
    ("n2049", lambda: _util.rand_cloud(3, 2049, 3, seed=14), 2048),
    ("streaming_ties_n9000", lambda: _util.tied_cloud(1, 9000, 3, levels=10, seed=11), 1200),
]


@pytest.mark.parametrize("cluster", [3, 2, 1, 0], ids=["pruned_single_cta", "cluster_async", "cluster_barrier", "single_cta"])
@pytest.mark.parametrize("name,make,S", FPS_CASES, ids=[c[0] for c in FPS_CASES])
def test_fps_bit_exact(dev, name, make, S, cluster):
    """Every kernel behind camli_furthest_point_sampling (Morton-bucketed single CTA with exact pruning / 8-CTA
    cluster with st.async exchange / with a cluster barrier per round / plain single CTA)."""
    from camliflow_b200 import native
    xyz = make().contiguous()
    old = native.lib().camli_fps_set_cluster_path(cluster)
    try:
        got = _csrc().furthest_point_sampling(xyz.to(dev), S)
        torch.cuda.synchronize()
    finally:
        native.lib().camli_fps_set_cluster_path(old)
    assert got.dtype == torch.int64 and got.shape == (xyz.shape[0], S)
    want = _util.oracle_fps(xyz.numpy(), S)
    assert np.array_equal(got.cpu().numpy(), want), "vs C oracle: %d mismatches" % (got.cpu().numpy() != want).sum()
    if _have_ref():
        ref = _util.ref_fps(xyz.to(dev), S)
        assert torch.equal(got, ref), "vs reference kernel: %d mismatches" % (got != ref).sum().item()


KNN_CASES = [
    # name, D, n, m, k, factory(B, N, D, seed)
    ("selftest_recipe", 3, 2048, 8192, 16, _util.rand_cloud),
    ("enc_4096x8192_k16", 3, 4096, 8192, 16, _util.rand_cloud),
    ("self_2048_k32", 3, 2048, 2048, 32, _util.rand_cloud),
    ("interp_8192x2048_k3", 3, 8192, 2048, 3, _util.rand_cloud),
    ("pool_256x512_k3", 3, 256, 512, 3, _util.rand_cloud),
    ("clfm2d_8160x2048_k1", 2, 8160, 2048, 1, _util.rand_cloud),
    ("clfm2d_34560x4096_k1", 2, 34560, 4096, 1, _util.rand_cloud),
    ("k64", 3, 500, 1000, 64, _util.rand_cloud),
    ("k33", 3, 300, 777, 33, _util.rand_cloud),
    ("k5_odd", 3, 37, 101, 5, _util.rand_cloud),
    ("m_lt_k", 3, 10, 7, 16, _util.rand_cloud),
    ("m_lt_k_64", 3, 10, 40, 64, _util.rand_cloud),
    ("m4_lt_k", 3, 5, 4, 6, _util.rand_cloud),
    ("ties3d_k16", 3, 1024, 2048, 16, _util.tied_cloud),
    ("ties3d_k32", 3, 512, 1500, 32, _util.tied_cloud),
    ("ties3d_k64", 3, 200, 900, 64, _util.tied_cloud),
    ("ties3d_k3", 3, 2048, 1024, 3, _util.tied_cloud),
    ("ties2d_k1", 2, 3000, 2048, 1, _util.tied_cloud),
    ("ties2d_k4", 2, 1000, 555, 4, _util.tied_cloud),
]


@pytest.mark.parametrize("name,D,n,m,k,make", KNN_CASES, ids=[c[0] for c in KNN_CASES])
def test_knn_bit_exact(dev, name, D, n, m, k, make):
    B = 2
    inp = make(B, m, D, seed=20)
    qry = make(B, n, D, seed=21)
    got = _csrc().k_nearest_neighbor(inp.to(dev), qry.to(dev), k)
    assert got.dtype == torch.int64 and got.shape == (B, n, k)
    want = _util.oracle_knn(inp.numpy(), qry.numpy(), k)
    assert np.array_equal(got.cpu().numpy(), want), "vs C oracle: %d mismatches" % (got.cpu().numpy() != want).sum()
    if _have_ref():
        ref = _util.ref_knn(inp.to(dev), qry.to(dev), k)
        assert torch.equal(got, ref), "vs reference kernel: %d mismatches" % (got != ref).sum().item()


def test_knn_self_query_and_channel_first(dev):
    # kNN(xyz1, xyz1) as camliraft_core.py:88 calls it: channel-first [B,3,N] tensors, d = 0 ties
    pc = _util.synthetic_pc(2, 2048, seed=30)
    pc[:, 100:140] = pc[:, 0:40]                      # exact duplicates
    cf = pc.transpose(1, 2).contiguous().to(dev)      # [B,3,N]
    got = _csrc().k_nearest_neighbor(cf, cf, 32)
    want = _util.oracle_knn(pc.numpy(), pc.numpy(), 32)
    assert np.array_equal(got.cpu().numpy(), want)
    got_cl = _csrc().k_nearest_neighbor(pc.to(dev), pc.to(dev), 32)
    assert torch.equal(got, got_cl)


CORR_CASES = [
    # name, B, C, H, W, md
    ("l5_192x9x15", 2, 192, 9, 15, 4),
    ("l4_128x18x30", 2, 128, 18, 30, 4),
    ("l3_96x36x60", 1, 96, 36, 60, 4),
    ("l2_64x72x120", 1, 64, 72, 120, 4),
    ("l1_32x144x240", 1, 32, 144, 240, 4),
    ("w33_edge", 1, 32, 5, 33, 4),
    ("generic_c20_d3", 2, 20, 11, 17, 3),
    ("generic_c48_d4", 1, 48, 10, 20, 4),
    ("generic_d1", 1, 32, 8, 8, 1),
]


@pytest.mark.parametrize("name,B,C,H,W,md", CORR_CASES, ids=[c[0] for c in CORR_CASES])
def test_correlation_forward_backward(dev, name, B, C, H, W, md):
    from camliflow_b200.csrc import wrapper
    g = torch.Generator().manual_seed(0)
    in1 = torch.rand((B, H, W, C), generator=g)
    in2 = torch.rand((B, H, W, C), generator=g)
    go = torch.rand((B, (2 * md + 1) ** 2, H, W), generator=g)
    out = wrapper._correlation_forward_cuda(in1.to(dev), in2.to(dev), md)
    g1, g2 = wrapper._correlation_backward_cuda(go.to(dev), in1.to(dev), in2.to(dev), md)
    want = _util.oracle_corr_fwd(in1.numpy(), in2.numpy(), md)
    w1, w2 = _util.oracle_corr_bwd(go.numpy(), in1.numpy(), in2.numpy(), md)
    # correlation_test.cpp:82-89 tolerance (mean |diff| < 1e-6), plus a max bound
    for a, b, what in ((out, want, "fwd"), (g1, w1, "grad1"), (g2, w2, "grad2")):
        diff = np.abs(a.cpu().numpy() - b)
        assert diff.mean() < 1e-6 and diff.max() < 2e-5, (what, diff.mean(), diff.max())
    if _have_ref():
        r = _util.ref_corr_fwd(in1.to(dev), in2.to(dev), md)
        r1, r2 = _util.ref_corr_bwd(go.to(dev), in1.to(dev), in2.to(dev), md)
        for a, b, what in ((out, r, "fwd"), (g1, r1, "grad1"), (g2, r2, "grad2")):
            diff = (a - b).abs()
            assert diff.mean().item() < 1e-6 and diff.max().item() < 2e-5, (what, diff.mean().item())


def test_correlation2d_autograd_matches_reference_formula(dev):
    """correlation2d (NCHW in, autograd) against the reference's own Python formula
    (wrapper.py:41-50: pad, 81 shifted products, channel mean) evaluated by torch autograd."""
    csrc = _csrc()
    g = torch.Generator().manual_seed(1)
    a = torch.rand((2, 64, 18, 30), generator=g).to(dev).requires_grad_(True)
    b = torch.rand((2, 64, 18, 30), generator=g).to(dev).requires_grad_(True)
    go = torch.rand((2, 81, 18, 30), generator=g).to(dev)
    out = csrc.correlation2d(a, b, 4)
    out.backward(go)
    ga, gb = a.grad.clone(), b.grad.clone()
    a.grad = None
    b.grad = None
    bp = torch.nn.functional.pad(b, [4] * 4)
    want = torch.cat([(a * bp[:, :, i:i + 18, j:j + 30]).mean(1, keepdim=True)
                      for i in range(9) for j in range(9)], 1)
    want.backward(go)
    assert (out - want).abs().max().item() < 2e-6
    assert (ga - a.grad).abs().max().item() < 2e-6 and (gb - b.grad).abs().max().item() < 2e-6


def test_golden_fixture_from_reference_python(dev):
    """Same fixtures the CPU suite pins the oracle with (tests/golden/l0_reference_py.npz)."""
    import os
    G = np.load(os.path.join(_util.GOLDEN, "l0_reference_py.npz"))
    csrc = _csrc()
    pc = _util.synthetic_pc(2, 8192, seed=0)
    got = csrc.furthest_point_sampling(pc.to(dev), 4096).cpu().numpy()
    assert np.array_equal(got, G["fps_synth_2x8192_s4096"].astype(np.int64))
    inp2 = (_util.rand_cloud(1, 2048, 2, seed=2) * 100).to(dev)
    qry2 = (_util.rand_cloud(1, 4000, 2, seed=3) * 100).to(dev)
    got = csrc.k_nearest_neighbor(inp2, qry2, 1).cpu().numpy()
    assert (got != G["knn2d_rand_1x4000x2048_k1"].astype(np.int64)).sum() <= 2
