"""GPU parity of the fused hot-path kernels (called through the C-ABI via camliflow_b200.ops)
against the plain-PyTorch fp32 formulas of tests/torch_ref.py on the same CUDA tensors.
Index-producing parts are bit-exact by construction (they share knn_search.cuh with the k-NN
kernel, itself bit-exact vs the reference kernels); floating-point results are compared with the
tolerances written below."""
import pytest
import torch

from tests import torch_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def _ops():
    from camliflow_b200 import ops
    return ops


def _knn(a, q, k):
    from camliflow_b200.csrc import k_nearest_neighbor
    return k_nearest_neighbor(a, q, k)


def _cloud(B, N, dev, seed, scale=10.0):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(B, 3, N, generator=g) - 0.5) * scale).to(dev)


def _close(a, b, atol, rtol=1e-5, what=""):
    err = (a - b).abs().max().item()
    assert torch.allclose(a, b, atol=atol, rtol=rtol), "%s max|diff| %.3e" % (what, err)


@pytest.mark.parametrize("B,m,n,F,k", [(1, 2048, 8192, 3, 3), (2, 2048, 2048, 3, 3), (2, 500, 333, 70, 3),
                                       (1, 64, 50, 5, 8)])
def test_knn_interpolate(dev, B, m, n, F, k):
    xi, xq = _cloud(B, m, dev, 1), _cloud(B, n, dev, 2)
    feat = torch.randn(B, F, m, device=dev)
    out = _ops().knn_interpolate(xi, feat, xq, k)
    ref = R.knn_interpolate(xi, feat, xq, _knn(xi, xq, k))
    _close(out, ref, 1e-5, what="knn_interpolate")


def test_knn_interpolate_coincident_points(dev):
    """Queries that coincide with inputs hit the clamp(1e-8) branch."""
    xi = _cloud(1, 1024, dev, 3)
    feat = torch.randn(1, 3, 1024, device=dev)
    out = _ops().knn_interpolate(xi, feat, xi[:, :, :512].contiguous(), 3)
    ref = R.knn_interpolate(xi, feat, xi[:, :, :512], _knn(xi, xi[:, :, :512].contiguous(), 3))
    _close(out, ref, 1e-5, what="coincident")


@pytest.mark.parametrize("B,m,n", [(1, 2048, 2048), (2, 1000, 700)])
def test_backwarp_3d_and_prefix_levels(dev, B, m, n):
    xyz1, xyz2 = _cloud(B, m, dev, 4), _cloud(B, n, dev, 5)
    flow = torch.randn(B, 3, m, device=dev) * 0.3
    out = _ops().backwarp_3d(xyz1, xyz2, flow, 3)
    warped = xyz1 + flow
    ref = xyz2 + R.knn_interpolate(warped, -flow, xyz2, _knn(warped, xyz2, 3))
    _close(out, ref, 1e-5, what="backwarp_3d")
    # every coarser level of a prefix pyramid is a prefix of the finest level's result
    half = xyz2[:, :, :n // 2].contiguous()
    out_half = _ops().backwarp_3d(xyz1, half, flow, 3)
    assert torch.equal(out_half, out[:, :, :n // 2])


@pytest.mark.parametrize("B,C,H,W,N", [(1, 128, 68, 120, 2048), (2, 324, 20, 28, 300), (1, 7, 5, 6, 50)])
def test_bilinear_sample(dev, B, C, H, W, N):
    g = torch.Generator().manual_seed(6)
    feat = torch.randn(B, C, H, W, generator=g).to(dev)
    uv = torch.stack([torch.rand(B, N, generator=g) * (W + 3) - 2, torch.rand(B, N, generator=g) * (H + 3) - 2], 1).to(dev)
    uv[:, :, :4] = torch.tensor([[0.0, W - 1.0, 3.0, W - 1.0], [0.0, H - 1.0, H - 1.0, 2.5]], device=dev)   # exact corners
    for src in (feat, feat.contiguous(memory_format=torch.channels_last)):
        out = _ops().bilinear_sample(src, uv)
        _close(out, R.bilinear_sample(feat, uv), 1e-5, what="bilinear_sample")


@pytest.mark.parametrize("B,C,H,W", [(1, 256, 68, 120), (2, 64, 17, 30), (1, 32, 9, 11)])
def test_corr2d_build_and_lookup(dev, B, C, H, W):
    g = torch.Generator().manual_seed(7)
    f1, f2 = torch.randn(B, C, H, W, generator=g).to(dev), torch.randn(B, C, H, W, generator=g).to(dev)
    L = 4 if min(H, W) >= 16 else 3
    pyr = _ops().corr2d_build(f1, f2, L)
    ref_pyr = R.corr2d_build(f1, f2, L)
    for a, b in zip(pyr, ref_pyr):
        assert a.shape == b.shape
        _close(a, b, 2e-5, what="corr2d_build level")
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    coords = (torch.stack([xs, ys], 0)[None].expand(B, 2, H, W) + torch.randn(B, 2, H, W, generator=g) * 6).to(dev)
    coords[:, :, 0, 0] = torch.tensor([0.0, 0.0], device=dev)          # exactly on a grid corner
    coords[:, :, 0, 1] = torch.tensor([-30.0, 500.0], device=dev)       # far outside
    coords[:, :, 1, 0] = torch.tensor([W - 1.0, H - 1.0], device=dev)
    ref = R.corr2d_lookup(ref_pyr, coords, 4)
    for cl in (True, False):
        out = _ops().corr2d_lookup(ref_pyr, coords, 4, channels_last=cl)
        assert out.shape == ref.shape
        # grid_sample re-derives every tap position through a normalise/un-normalise round trip
        # (|dx| ~ 1e-5 px at W=120), ours uses one exact fractional offset per pixel and level
        _close(out, ref, 2e-4, rtol=1e-4, what="corr2d_lookup cl=%s" % cl)


def test_corr3d_pool_and_lookup(dev):
    B, n1 = 2, 512
    g = torch.Generator().manual_seed(8)
    xyz1 = _cloud(B, n1, dev, 9)
    full = _cloud(B, 512, dev, 10)
    xyzs2 = [full[:, :, :n] for n in (512, 256, 128, 64)]              # prefix pyramid (strided views)
    vol = torch.randn(B, n1, 512, generator=g).to(dev)
    pyr_ref = [vol]
    for i in range(1, 4):
        pyr_ref.append(R.corr3d_pool(pyr_ref[-1], _knn(xyzs2[i - 1], xyzs2[i], 3)))
    feat1 = torch.randn(B, 32, n1, generator=g).to(dev)
    feat2 = torch.randn(B, 32, 512, generator=g).to(dev)
    pyr = _ops().corr3d_build(feat1, feat2, xyzs2, 3)
    _close(pyr[0], torch.bmm(feat1.transpose(1, 2), feat2) / 32, 1e-5, what="corr3d volume")
    p = pyr[0]
    for i in range(1, 4):
        p = R.corr3d_pool(p, _knn(xyzs2[i - 1], xyzs2[i], 3))
        _close(pyr[i], p, 1e-5, what="corr3d pool %d" % i)
    W1, b1 = torch.randn(32, 4, generator=g).to(dev) * 0.5, torch.randn(32, generator=g).to(dev) * 0.1
    W2, b2 = torch.randn(32, 32, generator=g).to(dev) * 0.2, torch.randn(32, generator=g).to(dev) * 0.1
    out = _ops().corr3d_lookup_rows(xyz1, xyzs2, pyr_ref, W1, b1, W2, b2)
    idxs = [_knn(x.contiguous(), xyz1, 16) for x in xyzs2]
    ref = R.corr3d_lookup(xyz1, xyzs2, pyr_ref, idxs, W1, b1, W2, b2)
    _close(out.transpose(1, 2), ref, 2e-4, rtol=1e-4, what="corr3d_lookup")


@pytest.mark.parametrize("B,N,S,K,k,O", [(1, 2048, 2048, 32, 32, 128), (2, 600, 600, 32, 16, 125), (1, 512, 512, 32, 4, 128),
                                         (1, 300, 100, 16, 16, 16)])
def test_pointconv_dw(dev, B, N, S, K, k, O):
    from camliflow_b200.mlp import MLP2d
    g = torch.Generator().manual_seed(11)
    xyz = _cloud(B, N, dev, 12)
    centre = xyz[:, :, :S].contiguous()
    idx = _knn(xyz, centre, K)
    wn = MLP2d(3, [8, 32, O], act="relu").to(dev)
    params = []
    for c in wn.convs:
        c.conv_fn.weight.data = torch.randn(c.conv_fn.weight.shape, generator=g).to(dev) * 0.4
        c.conv_fn.bias.data = torch.randn(c.conv_fn.bias.shape, generator=g).to(dev) * 0.2
        params += [c.conv_fn.weight.data.flatten(1), c.conv_fn.bias.data]
    with torch.no_grad():
        wc = _ops().pointconv_dw_weights(xyz, centre, idx, k, wn)
    ref_w = R.pointconv_dw_weights(xyz, centre, idx[:, :, :k], params)
    _close(wc, ref_w, 1e-4, rtol=1e-4, what="dw weights")
    feat = torch.randn(B, O, N, generator=g).to(dev)
    out = _ops().pointconv_dw_gather_max(_ops().rows_of(feat), ref_w, idx, k)
    ref = R.pointconv_dw_gather_max(feat, ref_w, idx[:, :, :k])
    assert torch.equal(out.transpose(1, 2), ref)      # product + max: no rounding freedom


@pytest.mark.parametrize("B,H,W,N,C", [(1, 68, 120, 2048, 128), (2, 9, 13, 100, 40)])
def test_clfm_interp(dev, B, H, W, N, C):
    from camliflow_b200.mlp import Conv2dNormRelu
    g = torch.Generator().manual_seed(13)
    uv = torch.stack([torch.rand(B, N, generator=g) * (W - 1), torch.rand(B, N, generator=g) * (H - 1)], 1).to(dev)
    feat3d = torch.randn(B, C, N, generator=g).to(dev)
    sn = torch.nn.Sequential(Conv2dNormRelu(3, 16), Conv2dNormRelu(16, C, act="sigmoid")).to(dev)
    nn_idx = _ops().nearest_point_2d(uv, H, W)
    with torch.no_grad():
        out = _ops().clfm_interp(uv, nn_idx, _ops().rows_of(feat3d), sn, H, W)
        ref = R.clfm_interp(uv, nn_idx, feat3d, sn[0].conv_fn.weight.flatten(1), sn[0].conv_fn.bias,
                            sn[1].conv_fn.weight.flatten(1), sn[1].conv_fn.bias, H, W)
    assert out.shape == ref.shape
    _close(out, ref, 1e-5, what="clfm_interp")


@pytest.mark.parametrize("B,N,S,C,k,act", [(1, 8192, 4096, 96, 16, "leaky_relu"), (2, 4096, 2048, 128, 16, "leaky_relu"),
                                           (1, 500, 200, 13, 9, "relu"), (1, 300, 300, 195, 16, None)])
def test_pointconv_group(dev, B, N, S, C, k, act):
    from camliflow_b200.mlp import MLP2d
    g = torch.Generator().manual_seed(14)
    xyz = _cloud(B, N, dev, 15)
    centre = xyz[:, :, :S].contiguous()
    feat = torch.randn(B, C, N, generator=g).to(dev)
    idx = _knn(xyz, centre, k)
    wn = MLP2d(3, [8, 16], act=act).to(dev)
    slope = {"relu": 0.0, "leaky_relu": 0.1, None: 1.0}[act]
    with torch.no_grad():
        out = _ops().pointconv_group(_ops().rows_of(torch.cat([xyz, feat], 1)), centre, idx, k, wn, slope)
        ref = R.pointconv_group(xyz, feat, centre, idx, wn.convs[0].conv_fn.weight.flatten(1), wn.convs[0].conv_fn.bias,
                                wn.convs[1].conv_fn.weight.flatten(1), wn.convs[1].conv_fn.bias, slope)
    _close(out, ref, 1e-4, rtol=1e-4, what="pointconv_group")


@pytest.mark.parametrize("B,M,N,K", [(1, 8160, 8160, 256), (2, 300, 500, 64), (1, 2048, 2048, 128), (3, 128, 128, 32),
                                     (1, 129, 257, 96)])
def test_allpairs_tcgen05(dev, B, M, N, K):
    """3xTF32 tensor-core product vs an fp64 reference: fp32-level accuracy is the bar (the reference
    computes this product in fp32), so the error must be no worse than ~2x a true fp32 SGEMM's."""
    g = torch.Generator().manual_seed(16)
    a = torch.randn(B, M, K, generator=g).to(dev)
    b = torch.randn(B, N, K, generator=g).to(dev)
    scale = 1.0 / K ** 0.5
    out = _ops().allpairs(a, b, scale)
    torch.cuda.synchronize()
    ref64 = torch.bmm(a.double(), b.double().transpose(1, 2)) * scale
    sgemm = torch.bmm(a, b.transpose(1, 2)) * scale
    err = (out.double() - ref64).abs().max().item()
    err_sgemm = (sgemm.double() - ref64).abs().max().item()
    print("allpairs %s: max err %.3e (fp32 SGEMM %.3e)" % ((B, M, N, K), err, err_sgemm))
    assert err <= max(4 * err_sgemm, 2e-6), (err, err_sgemm)


@pytest.mark.parametrize("B,P,C", [(1, 8160, 324), (2, 2048, 128), (1, 100, 20)])
def test_sk_fusion_tail(dev, B, P, C):
    g = torch.Generator().manual_seed(17)
    a, b = torch.randn(B, P, C, generator=g).to(dev), torch.randn(B, P, C, generator=g).to(dev)
    w_mid = (torch.randn(C // 2, C, generator=g) * 0.2).to(dev)
    w_out = (torch.randn(2 * C, C // 2, generator=g) * 0.2).to(dev)
    for slope in (1.0, 0.1):
        out = _ops().sk_fusion_tail(a, b, slope, w_mid, w_out)
        _close(out, R.sk_fusion_tail(a, b, slope, w_mid, w_out), 2e-5, what="sk_fusion_tail")
        # the same into a channel slice of a wider channel-last buffer (camli_sk_fusion_tail_strided)
        wide = torch.full((B, P, C + 24), 7.0, device=dev)
        got = _ops().sk_fusion_tail(a, b, slope, w_mid, w_out, out=wide[..., 8:8 + C])
        assert got.data_ptr() == wide[..., 8:8 + C].data_ptr() and torch.equal(wide[..., 8:8 + C], out)
        assert bool((wide[..., :8] == 7.0).all()) and bool((wide[..., 8 + C:] == 7.0).all())


@pytest.mark.parametrize("channels_last", [True, False])
def test_gru_gates(dev, channels_last):
    g = torch.Generator().manual_seed(18)
    B, H, X, hh, ww = 2, 128, 256, 17, 30
    zr, h, x, q = (torch.randn(B, c, hh, ww, generator=g).to(dev) for c in (2 * H, H, X, H))
    if channels_last:
        zr, h, x, q = (t.contiguous(memory_format=torch.channels_last) for t in (zr, h, x, q))
    z, rhx = _ops().gru_gate(zr, h, x)
    _close(z, torch.sigmoid(zr[:, :H]), 1e-6, what="gru z")
    _close(rhx, torch.cat([torch.sigmoid(zr[:, H:]) * h, x], 1), 1e-6, what="gru rhx")
    q[0, 0, 0, 0] = float("nan")
    q[0, 1, 0, 0] = float("inf")
    out = _ops().gru_update(z, h, q, fix_nonfinite=True)
    _close(out, torch.nan_to_num((1 - z) * h + z * torch.tanh(q)), 1e-6, what="gru update")


@pytest.mark.parametrize("channels_last", [True, False])
def test_encoder2d_fused_epilogues(dev, channels_last):
    """Inference path of Encoder2D (BatchNorm folded, conv+bias(+residual)+ReLU as single cuDNN calls)
    against the module-by-module path autograd uses (conv, BatchNorm, add, ReLU)."""
    from camliflow_b200.raft_core import Encoder2D
    g = torch.Generator().manual_seed(21)
    enc = Encoder2D().eval()
    for m in enc.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    enc = enc.to(dev)
    x = torch.randn(2, 3, 96, 128, generator=g).to(dev)
    if channels_last:
        enc, x = enc.to(memory_format=torch.channels_last), x.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        fused = enc(x)
    with torch.enable_grad():
        plain = enc(x).detach()
    scale = plain.abs().max().item()
    _close(fused, plain, 2e-5 * max(scale, 1.0), rtol=1e-4, what="Encoder2D fused epilogues")


def _tc_weight(w_oihw, bias):
    """Weight [O,I,kh,kw] -> the OHWI hi/lo split camli_conv_gemm consumes."""
    O = w_oihw.shape[0]
    return _ops().tc_weight([w_oihw], lambda: (w_oihw.permute(0, 2, 3, 1).reshape(O, -1), bias))


CONV_GEMM_CASES = [
    # B, H, W, Cin, Cout, kh, kw, act, residual, tile_n
    (1, 1, 2048, 384, 128, 1, 1, None, False, 0),            # PointConvDW mlp of the 3-D GRU
    (1, 1, 2048, 144, 125, 1, 1, "leaky_relu", False, 0),    # ragged Cout, K tail (144 = 4.5 k-blocks)
    (2, 1, 300, 36, 3, 1, 1, None, False, 0),                # tiny head
    (1, 1, 4096, 1584, 96, 1, 1, "leaky_relu", False, 0),    # PointConv Linear of the encoder
    (1, 68, 120, 256, 192, 3, 3, "relu", False, 0),          # MotionEncoder2D.conv_c2
    (1, 68, 120, 384, 256, 1, 5, None, False, 0),            # merged z|r convolution of the ConvGRU
    (1, 68, 120, 384, 128, 5, 1, "tanh", False, 64),
    (2, 30, 44, 64, 64, 3, 3, "relu", True, 0),              # bottleneck tail: conv + residual + ReLU
    (1, 17, 23, 8, 40, 7, 7, "sigmoid", False, 32),          # Cin below one k-block, wide window
    (1, 136, 240, 64, 256, 1, 1, "relu", True, 128),         # encoder 1x1, widest accumulator tile
    (1, 20, 36, 128, 256, 3, 3, None, False, 128),
    # 96-column tiles (the automatic width of 96 / 192 / 288-channel layers): 64 | 32 column split of the epilogue warps
    (1, 68, 120, 256, 192, 3, 3, "relu", True, 96),
    (2, 30, 44, 64, 288, 1, 1, "leaky_relu", False, 0),
    (1, 17, 23, 40, 100, 3, 3, "tanh", True, 96),           # ragged: a 96-column tile and one with 4 live columns
    (1, 1, 300, 128, 96, 1, 1, None, False, 0),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,kh,kw,act,res,tile_n", CONV_GEMM_CASES)
def test_conv_gemm_tcgen05(dev, B, H, W, Cin, Cout, kh, kw, act, res, tile_n):
    """3xTF32 implicit-GEMM convolution / linear layer against an fp64 convolution: the error must stay at
    the level of an fp32 SGEMM (a plain TF32 product is ~1000x worse)."""
    g = torch.Generator().manual_seed(31)
    x = torch.randn(B, Cin, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, kh, kw, generator=g) / (Cin * kh * kw) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    r = torch.randn(B, Cout, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last) if res else None
    w_hi, w_lo, bias = _tc_weight(w, b)
    out = _ops().conv_gemm(x.permute(0, 2, 3, 1), w_hi, w_lo, kh, kw, bias, act, 0.1,
                           None if r is None else r.permute(0, 2, 3, 1), tile_n=tile_n).permute(0, 3, 1, 2)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=(kh // 2, kw // 2))
    if r is not None:
        ref = ref + r.double()
    ref = {None: lambda v: v, "relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid,
           "leaky_relu": lambda v: torch.nn.functional.leaky_relu(v, 0.1)}[act](ref)
    err = (out.double() - ref).abs().max().item()
    fp32 = torch.nn.functional.conv2d(x, w, b, padding=(kh // 2, kw // 2))
    print("conv_gemm %s: max err %.3e" % ((B, H, W, Cin, Cout, kh, kw), err))
    assert err <= 2e-5, err
    assert out.shape == fp32.shape


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,act,res", [
    (2, 136, 240, 128, 128, 3, "relu", False),      # layer2.0 conv2 of the ResNet encoder (stride 2 on the 3x3)
    (1, 136, 240, 256, 512, 1, None, False),        # layer2.0 downsample
    (2, 37, 61, 64, 96, 3, "relu", True),           # odd input grid: ceil(H/2) x ceil(W/2) outputs, ragged tiles
    (1, 20, 33, 32, 40, 5, "relu_fix", False),
])
def test_conv_gemm_stride2(dev, B, H, W, Cin, Cout, k, act, res):
    """Stride-2 convolution through the TMA element-stride tensor map against an fp64 convolution."""
    g = torch.Generator().manual_seed(33)
    x = torch.randn(B, Cin, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=2, padding=k // 2)
    r = None
    if res:
        r = torch.randn(ref.shape, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
        ref = ref + r.double()
    ref = torch.relu(ref) if act else ref
    w_hi, w_lo, bias = _tc_weight(w, b)
    out = _ops().conv_gemm(x.permute(0, 2, 3, 1), w_hi, w_lo, k, k, bias, act, 0.1,
                           None if r is None else r.permute(0, 2, 3, 1), stride=2).permute(0, 3, 1, 2)
    assert out.shape == ref.shape, (out.shape, ref.shape)
    err = (out.double() - ref).abs().max().item()
    print("conv_gemm stride 2 %s: max err %.3e" % ((B, H, W, Cin, Cout, k), err))
    assert err <= 2e-5, err


def test_conv_gemm_nan_to_num_epilogue(dev):
    """act "relu_fix" / "none_fix" = torch.nan_to_num(act(conv)) (models/raft_core.py:164,180): outputs reached by a
    NaN activation become 0.  (An INFINITE activation turns into NaN inside the tf32 hi/lo split -- inf - inf -- so it
    also ends as 0 where the reference would write FLT_MAX; neither value means anything downstream.)"""
    g = torch.Generator().manual_seed(34)
    x = torch.randn(1, 16, 12, 32, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    x[0, 0, 3, 5] = float("nan")
    x[0, 1, 7, 9] = float("nan")
    w = (torch.randn(40, 16, 3, 3, generator=g) / 12).to(dev)
    w_hi, w_lo, _ = _tc_weight(w, None)
    for act, fn in (("relu_fix", torch.relu), ("none_fix", lambda v: v)):
        out = _ops().conv_gemm(x.permute(0, 2, 3, 1), w_hi, w_lo, 3, 3, None, act).permute(0, 3, 1, 2)
        ref = torch.nan_to_num(fn(torch.nn.functional.conv2d(x, w, padding=1)))
        assert torch.isfinite(out).all()
        finite = torch.isfinite(torch.nn.functional.conv2d(x, w, padding=1))
        assert torch.allclose(out[finite], ref[finite], atol=2e-5)
        assert torch.equal(out[~finite] == 0, ref[~finite] == 0)         # nan -> 0, +-inf -> +-FLT_MAX (or 0 after relu)


@pytest.mark.parametrize("B,H,W", [(2, 544, 960), (1, 160, 224), (1, 75, 133)])
def test_stem_conv_pool(dev, B, H, W):
    """Fused ResNet stem (7x7 s2 conv + bias + ReLU + 3x3 s2 max-pool, fp32 FMA) against torch in fp64."""
    g = torch.Generator().manual_seed(35)
    x = torch.randn(B, 3, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(64, 3, 7, 7, generator=g) / 12).to(dev)
    b = torch.randn(64, generator=g).to(dev)
    out = _ops().stem_conv_pool(x.permute(0, 2, 3, 1), w.permute(0, 2, 3, 1).contiguous(), b).permute(0, 3, 1, 2)
    ref = torch.nn.functional.max_pool2d(torch.relu(torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=2, padding=3)), 3, 2, 1)
    assert out.shape == ref.shape, (out.shape, ref.shape)
    err = (out.double() - ref).abs().max().item()
    print("stem %s: max err %.3e" % ((B, H, W), err))
    assert err <= 2e-5, err


def test_conv_gemm_channel_slices(dev):
    """Input read from, and output written into, channel slices of wider channel-last buffers (how the
    update block avoids torch.cat)."""
    g = torch.Generator().manual_seed(32)
    wide_in = torch.randn(1, 20, 28, 96, generator=g).to(dev)
    wide_out = torch.zeros(1, 20, 28, 80, device=dev)
    w = (torch.randn(48, 64, 3, 3, generator=g) / 24.0).to(dev)
    w_hi, w_lo, _ = _tc_weight(w, None)
    x = wide_in[..., 32:]                       # 64 channels at offset 32
    _ops().conv_gemm(x, w_hi, w_lo, 3, 3, None, "relu", out=wide_out[..., 16:64])
    ref = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), padding=1)).permute(0, 2, 3, 1)
    assert (wide_out[..., 16:64].double() - ref).abs().max().item() <= 2e-5
    assert wide_out[..., :16].abs().max().item() == 0 and wide_out[..., 64:].abs().max().item() == 0


@pytest.mark.parametrize("B,H,W,Cin,Cout,kh,kw,act", [(1, 68, 120, 256, 2, 3, 3, None), (2, 1, 2048, 64, 3, 1, 1, None),
                                                      (1, 9, 11, 8, 4, 5, 3, "relu"), (1, 5, 7, 12, 1, 1, 1, "sigmoid")])
def test_conv_small_n(dev, B, H, W, Cin, Cout, kh, kw, act):
    g = torch.Generator().manual_seed(33)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, kh, kw, generator=g) / (Cin * kh * kw) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    out = _ops().conv_small_n(x, w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous(), kh, kw, b, act)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=(kh // 2, kw // 2))
    ref = {None: lambda v: v, "relu": torch.relu, "sigmoid": torch.sigmoid}[act](ref).permute(0, 2, 3, 1)
    assert (out.double() - ref).abs().max().item() <= 5e-6


def test_gru3d_fused_matches_layerwise(dev):
    """GRU3D with the merged z|r point convolution + gate kernels against its layer-by-layer form."""
    from camliflow_b200 import tc
    from camliflow_b200.camliraft_l_core import GRU3D
    g = torch.Generator().manual_seed(34)
    gru = GRU3D(input_dim=256, hidden_dim=128)
    for p in gru.parameters():
        p.data.copy_(torch.randn(p.shape, generator=g) * 0.1)
    gru = gru.to(dev).eval()
    xyz = _cloud(1, 700, dev, 35)
    h, x = torch.randn(1, 700, 128, generator=g).to(dev), torch.randn(1, 700, 256, generator=g).to(dev)
    nbr = _knn(xyz, xyz, 32)
    with torch.no_grad():
        fused = gru.forward_rows(xyz, h, x, nbr, {})
        tc.ENABLED = False
        try:
            plain = gru.forward_rows(xyz, h, x, nbr, {})
        finally:
            tc.ENABLED = True
    _close(fused, plain, 2e-5, rtol=1e-4, what="GRU3D fused")


def test_gru2d_split_matches_plain(dev):
    """ConvGRU with the context contribution precomputed (forward_split) against the plain module."""
    from camliflow_b200.raft_core import GRU2D
    g = torch.Generator().manual_seed(36)
    gru = GRU2D(hidden_dim=128, input_dim=256)
    for p in gru.parameters():
        p.data.copy_(torch.randn(p.shape, generator=g) * 0.03)
    gru = gru.to(dev).eval().to(memory_format=torch.channels_last)
    mk = lambda c: torch.randn(1, c, 20, 28, generator=g).to(dev).contiguous(memory_format=torch.channels_last)  # noqa: E731
    h, xs, xd = torch.tanh(mk(128)), torch.relu(mk(128)), mk(128)
    with torch.no_grad():
        cache = {}
        a = gru.forward_split(h, xs, xd, cache)
        a2 = gru.forward_split(h, xs, xd, cache)           # second call reuses the cached context terms
    with torch.enable_grad():
        b = gru(h, torch.cat([xs, xd], 1)).detach()        # module-by-module torch path
    assert torch.equal(a, a2)
    _close(a, b, 2e-5, rtol=1e-4, what="GRU2D split")


@pytest.mark.parametrize("B,H,W,Cin,Cout,kh,kw,act", [(1, 68, 120, 2, 128, 7, 7, "relu"), (2, 13, 37, 3, 40, 3, 5, None),
                                                      (1, 9, 33, 4, 32, 1, 1, "sigmoid")])
def test_conv_small_cin(dev, B, H, W, Cin, Cout, kh, kw, act):
    g = torch.Generator().manual_seed(37)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, kh, kw, generator=g) / (Cin * kh * kw) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    out = _ops().conv_small_cin(x, w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous(), kh, kw, b, act)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=(kh // 2, kw // 2))
    ref = {None: lambda v: v, "relu": torch.relu, "sigmoid": torch.sigmoid}[act](ref).permute(0, 2, 3, 1)
    assert (out.double() - ref).abs().max().item() <= 5e-6


def test_pointconv_dw_weights_tensor_core_route(dev):
    """Inference route of the PointConvDW WeightNet (hidden-layer kernel + tensor-core output layer) against the
    formula; the autograd route (single fused kernel) against the same."""
    from camliflow_b200.mlp import MLP2d
    g = torch.Generator().manual_seed(38)
    xyz = _cloud(2, 700, dev, 39)
    nbr = _knn(xyz, xyz, 32)
    wn = MLP2d(3, [8, 32, 128], act="relu").to(dev)
    fp = []
    for c in wn.convs:
        fp += [c.conv_fn.weight.flatten(1), c.conv_fn.bias]
    for k in (32, 16, 4):
        with torch.no_grad():
            fast = _ops().pointconv_dw_weights(xyz, xyz, nbr, k, wn)
            ref = R.pointconv_dw_weights(xyz, xyz, nbr[:, :, :k], fp)
        slow = _ops().pointconv_dw_weights(xyz, xyz, nbr, k, wn).detach()       # parameters require grad: fused kernel
        _close(fast, ref, 2e-5, rtol=1e-4, what="dw_weights tensor-core route k=%d" % k)
        _close(slow, ref, 2e-5, rtol=1e-4, what="dw_weights fused kernel k=%d" % k)


def test_dense_three_nn_kitti_size(dev):
    """the KITTI-submission densification, 375 x 1242 = 465 750 queries against 8192 points
    (kitti_submission.py:89-93), against the formula with the kernel's own neighbour indices."""
    import time
    from camliflow_b200.utils import densify_flow_3d
    g = torch.Generator().manual_seed(40)
    H, W, f, cx, cy = 375, 1242, 721.5, 609.5, 172.8
    disp = (torch.rand(H, W, generator=g) * 70 + 5).to(dev)
    u, v, z = torch.rand(8192, generator=g) * (W - 1), torch.rand(8192, generator=g) * (H - 1), torch.rand(8192, generator=g) * 60 + 5
    pc1 = torch.stack([(u - cx) * z / f, (v - cy) * z / f, z], 0).to(dev)
    flow = (torch.randn(3, 8192, generator=g) * 0.3).to(dev)
    dense, got = densify_flow_3d(pc1, flow, disp, 0.54, f, cx, cy)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dense, got = densify_flow_3d(pc1, flow, disp, 0.54, f, cx, cy)
    torch.cuda.synchronize()
    print("dense three-NN, %d queries x 8192 points: %.2f ms" % (H * W, (time.perf_counter() - t0) * 1e3))
    idx = _knn(pc1[None], dense[None], 3)
    ref = R.knn_interpolate(pc1[None], flow[None], dense[None], idx)[0]
    _close(got, ref, 1e-5, what="dense three-NN")
