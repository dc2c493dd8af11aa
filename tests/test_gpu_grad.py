"""GPU parity of the training path: gradients of the fused operators (kernel forward + hand-written or
recompute backward, camliflow_b200/grad.py) against torch autograd through the plain-PyTorch formulas of
tests/torch_ref.py, per operator and through a whole CamLiRAFT training step (forward, sequence losses,
backward).  Tolerances are written at each comparison."""
import contextlib

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


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _own_decisions(gy, y_kernel, y_ref_lin, act, slope=0.1):
    """(reference output to differentiate, upstream gradient) such that a piecewise-linear activation uses the KERNEL's
    decisions: an fp32 output within rounding of zero may land on the other side of the kink than the fp64 reference's,
    which flips whole gradient entries (the more outputs, the likelier) and says nothing about the GEMMs under test."""
    if act == "relu":
        return y_ref_lin, gy * (y_kernel > 0).double()
    if act == "leaky_relu":
        return y_ref_lin, gy * torch.where(y_kernel > 0, 1.0, slope).double()
    if act == "tanh":
        return torch.tanh(y_ref_lin), gy
    if act == "sigmoid":
        return torch.sigmoid(y_ref_lin), gy
    return y_ref_lin, gy


def test_dw_gather_max_backward(dev):
    g = torch.Generator().manual_seed(41)
    B, N, S, K, k, O = 2, 600, 500, 32, 16, 125
    xyz = ((torch.rand(B, 3, N, generator=g) - 0.5) * 10).to(dev)
    idx = _knn(xyz, xyz[:, :, :S].contiguous(), K)
    feat = torch.randn(B, N, O, generator=g).to(dev).requires_grad_(True)
    w = torch.rand(B, S, k, O, generator=g).to(dev).requires_grad_(True)
    gout = torch.randn(B, S, O, generator=g).to(dev)
    out = _ops().pointconv_dw_gather_max(feat, w, idx, k)
    gf, gw = torch.autograd.grad(out, [feat, w], gout)
    ref = R.pointconv_dw_gather_max(feat.transpose(1, 2), w, idx[:, :, :k]).transpose(1, 2)
    rf, rw = torch.autograd.grad(ref, [feat, w], gout)
    assert _rel(out, ref) <= 1e-6
    assert _rel(gf, rf) <= 1e-5 and _rel(gw, rw) <= 1e-6      # g_feat is an atomic sum: order-dependent rounding


@pytest.mark.parametrize("B,H,W,Cin,Cout,kh,kw,dil,act,bias", [
    (1, 68, 120, 256, 192, 3, 3, 1, "relu", True),        # motion encoder conv_c2 at C2 size
    (2, 20, 36, 64, 126, 3, 3, 1, "relu", True),          # ragged rows (W % 32 != 0), C_out % 4 != 0 (padded by the doorway)
    (1, 24, 32, 147, 128, 3, 3, 1, "leaky_relu", True),   # C_in % 4 != 0 (PWC dense estimator): zero-padded channels
    (1, 17, 40, 128, 128, 1, 5, 1, "sigmoid", True),      # ConvGRU gate
    (1, 17, 40, 128, 128, 5, 1, 1, "tanh", True),
    (2, 24, 32, 32, 64, 3, 3, 2, "leaky_relu", False),    # dilated (PWC context network)
    (1, 1, 2048, 384, 256, 1, 1, 1, None, True),          # point-branch linear
    (1, 1, 520, 128, 64, 1, 1, 1, "leaky_relu", True),
])
def test_dense_layer_backward_kernels(dev, B, H, W, Cin, Cout, kh, kw, dil, act, bias):
    """The training doorway tc.conv_train -> grad.DenseFn (forward conv_gemm; backward transpose_split + conv_gemm on mirrored
    weights + conv_wgrad) against fp64 autograd through F.conv2d: output, data gradient, weight gradient, bias gradient."""
    import torch.nn.functional as F
    from camliflow_b200 import grad, tc
    g = torch.Generator().manual_seed(77)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev).requires_grad_(True)
    w = (torch.randn(Cout, Cin, kh, kw, generator=g) / (Cin * kh * kw) ** 0.5).to(dev).requires_grad_(True)
    b = (torch.randn(Cout, generator=g) * 0.1).to(dev).requires_grad_(True) if bias else None
    gy = torch.randn(B, H, W, Cout, generator=g).to(dev)
    grad.clear_dense_cache()
    y = tc.conv_train(x.permute(0, 3, 1, 2), w, b, (1, 1), (dil * (kh // 2), dil * (kw // 2)), (dil, dil), 1, act, 0.1)
    assert y is not None
    y = y.permute(0, 2, 3, 1)
    got = torch.autograd.grad(y, [x, w] + ([b] if bias else []), gy)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    yd = F.conv2d(xd.permute(0, 3, 1, 2), wd, bd, padding=(dil * (kh // 2), dil * (kw // 2)), dilation=dil)
    ylin = yd.permute(0, 2, 3, 1)
    yd = {None: lambda v: v, "relu": torch.relu, "leaky_relu": lambda v: F.leaky_relu(v, 0.1), "tanh": torch.tanh,
          "sigmoid": torch.sigmoid}[act](ylin)
    out, up = _own_decisions(gy.double(), y.detach(), ylin, act)
    ref = torch.autograd.grad(out, [xd, wd] + ([bd] if bias else []), up)
    names = ["dx", "dw", "db"]
    print("dense %s: y %.2e %s" % ((B, H, W, Cin, Cout, kh, kw, dil, act), _rel(y.double(), yd),
                                    " ".join("%s %.2e" % (n, _rel(a.double(), r)) for n, a, r in zip(names, got, ref))))
    assert _rel(y.double(), yd) <= 1e-5
    for n, a, r in zip(names, got, ref):
        assert a.shape == r.shape
        assert _rel(a.double(), r) <= 2e-5, n            # fp32-level: 3xTF32 products, fp32 accumulation


@pytest.mark.parametrize("B,H,W,Cin,Cout,kh,kw,act", [(1, 68, 120, 256, 192, 3, 3, "relu"), (1, 1, 2048, 384, 256, 1, 1, None)])
def test_dense_layer_single_pass_mode(dev, B, H, W, Cin, Cout, kh, kw, act):
    """passes = 1 (the mode picked under bf16 autocast): one tf32 product per element in forward, data gradient and weight
    gradient -- tf32-operand accuracy (10-bit mantissas, fp32 accumulation), far inside what bf16 layers deliver."""
    import torch.nn.functional as F
    from camliflow_b200 import grad
    g = torch.Generator().manual_seed(78)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev).requires_grad_(True)
    w = (torch.randn(Cout, Cin, kh, kw, generator=g) / (Cin * kh * kw) ** 0.5).to(dev).requires_grad_(True)
    b = (torch.randn(Cout, generator=g) * 0.1).to(dev).requires_grad_(True)
    gy = torch.randn(B, H, W, Cout, generator=g).to(dev)
    grad.clear_dense_cache()
    y = grad.DenseFn.apply(x, w, b, act, 0.1, 1, 1)
    got = torch.autograd.grad(y, [x, w, b], gy)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    yd = F.conv2d(xd.permute(0, 3, 1, 2), wd, bd, padding=(kh // 2, kw // 2))
    yd = (torch.relu(yd) if act == "relu" else yd).permute(0, 2, 3, 1)
    # gradients against the reference with the kernel's OWN ReLU decisions (an output within rounding of zero may land on the
    # other side in any reduced-precision forward; that flips whole gradient entries and says nothing about the GEMMs)
    ylin = F.conv2d(xd.permute(0, 3, 1, 2), wd, bd, padding=(kh // 2, kw // 2)).permute(0, 2, 3, 1)
    gmask = gy.double() * (y.detach() > 0).double() if act == "relu" else gy.double()
    ref = torch.autograd.grad(ylin, [xd, wd, bd], gmask)
    errs = [_rel(y.double(), yd)] + [_rel(a.double(), r) for a, r in zip(got, ref)]
    xb, wb = x.detach().bfloat16().float().double(), w.detach().bfloat16().float().double()       # what bf16 operands would give
    yb = F.conv2d(xb.permute(0, 3, 1, 2), wb, bd.detach(), padding=(kh // 2, kw // 2))
    yb = (torch.relu(yb) if act == "relu" else yb).permute(0, 2, 3, 1)
    print("single pass %s: y %.2e dx %.2e dw %.2e db %.2e (bf16-rounded operands: y %.2e)" %
          ((B, H, W, Cin, Cout, kh, kw, act), errs[0], errs[1], errs[2], errs[3], _rel(yb, yd)))
    # forward and data gradient: one tf32 product; weight gradient: bf16 operands (grad.BF16_WGRAD), fp32 accumulation
    assert max(errs[:2]) <= 2e-3 and errs[2] <= 8e-3 and errs[3] <= 2e-5
    assert errs[0] <= _rel(yb, yd)
    grad.BF16_WGRAD = False                              # the same with tf32 operands in the weight gradient
    try:
        grad.clear_dense_cache()
        y2 = grad.DenseFn.apply(x, w, b, act, 0.1, 1, 1)
        dw2 = torch.autograd.grad(y2, [w], gy)[0]
        assert _rel(dw2.double(), ref[1]) <= 2e-3
    finally:
        grad.BF16_WGRAD = True


@pytest.mark.parametrize("B,Hin,Win,Cin,Cout,k,bias", [(2, 68, 120, 128, 128, 3, False), (1, 135, 240, 256, 512, 1, False), (1, 34, 56, 64, 96, 3, True)])
def test_dense_layer_stride2_backward(dev, B, Hin, Win, Cin, Cout, k, bias):
    """Stride-2 layers of the encoders (raft_core.py:10-38) through the training doorway: strided forward, data gradient as the
    stride-1 convolution of the zero-upsampled gradient, weight gradient over the x-subsampled transposed copies."""
    import torch.nn.functional as F
    from camliflow_b200 import grad, tc
    g = torch.Generator().manual_seed(79)
    x = torch.randn(B, Hin, Win, Cin, generator=g).to(dev).requires_grad_(True)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev).requires_grad_(True)
    b = (torch.randn(Cout, generator=g) * 0.1).to(dev).requires_grad_(True) if bias else None
    grad.clear_dense_cache()
    y = tc.conv_train(x.permute(0, 3, 1, 2), w, b, (2, 2), (k // 2, k // 2), (1, 1), 1, "relu", 0.1)
    assert y is not None
    y = y.permute(0, 2, 3, 1)
    gy = torch.randn(y.shape, generator=g).to(dev)
    got = torch.autograd.grad(y, [x, w] + ([b] if bias else []), gy)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    ylin = F.conv2d(xd.permute(0, 3, 1, 2), wd, bd, stride=2, padding=k // 2).permute(0, 2, 3, 1)
    yd = torch.relu(ylin)
    out, up = _own_decisions(gy.double(), y.detach(), ylin, "relu")
    ref = torch.autograd.grad(out, [xd, wd] + ([bd] if bias else []), up)
    errs = [_rel(y.double(), yd)] + [_rel(a.double(), r) for a, r in zip(got, ref)]
    print("stride 2 %s:" % ((B, Hin, Win, Cin, Cout, k),), " ".join("%.2e" % e for e in errs))
    assert y.shape == yd.shape and max(errs) <= 2e-5


def test_dense_layers_route_through_the_kernels_under_autograd(dev):
    """nn.Conv2d / nn.Conv1d / nn.Linear applied through the tc doorways under autograd give the gradients of the library
    route (cuDNN / cuBLAS, strict fp32) -- and really take the kernel route (launch counter)."""
    import torch.nn as nn
    from camliflow_b200 import native, tc, grad
    g = torch.Generator().manual_seed(5)
    conv = nn.Conv2d(128, 256, 3, padding=1).to(dev)
    lin = nn.Linear(128, 64).to(dev)
    c1d = nn.Conv1d(64, 32, 1).to(dev)
    x = torch.randn(2, 128, 24, 40, generator=g).to(dev).requires_grad_(True)
    p = torch.randn(2, 600, 128, generator=g).to(dev).requires_grad_(True)

    def run(route):
        old = tc.TRAIN_DENSE
        tc.TRAIN_DENSE = route
        try:
            grad.clear_dense_cache()
            for m in (conv, lin, c1d):
                m.zero_grad()
            n0 = native.launch_count()
            y = tc.conv2d(x, conv, "relu")
            f = tc.linear(p, lin.weight, lin.bias, "leaky_relu")
            q = tc.module_train(c1d, f.transpose(1, 2))
            q = c1d(f.transpose(1, 2)) if q is None else q
            loss = y.square().mean() + q.square().mean()
            gx, gp = torch.autograd.grad(loss, [x, p], retain_graph=True)
            loss.backward()
            grads = [gx, gp] + [t.grad.clone() for m in (conv, lin, c1d) for t in m.parameters()]
            return float(loss), grads, native.launch_count() - n0
        finally:
            tc.TRAIN_DENSE = old

    l_lib, g_lib, n_lib = run("library")
    l_tc, g_tc, n_tc = run("tcgen05")
    assert n_lib == 0 and n_tc >= 15                     # forward + backward kernels of three layers went through the C ABI
    assert abs(l_lib - l_tc) <= 1e-6 * abs(l_lib)
    for a, b in zip(g_tc, g_lib):
        assert _rel(a, b) <= 2e-5


@pytest.mark.parametrize("B,C,H,W", [(1, 64, 24, 40), (2, 32, 17, 30)])
def test_corr2d_build_lookup_backward(dev, B, C, H, W):
    g = torch.Generator().manual_seed(42)
    f1 = torch.randn(B, C, H, W, generator=g).to(dev).requires_grad_(True)
    f2 = torch.randn(B, C, H, W, generator=g).to(dev).requires_grad_(True)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    coords = (torch.stack([xs, ys], 0)[None] + torch.randn(B, 2, H, W, generator=g) * 3).to(dev)
    coords[:, :, 0, 0] = torch.tensor([W + 20.0, -30.0], device=dev)           # a window entirely outside the map
    gout = torch.randn(B, 4 * 81, H, W, generator=g).to(dev)
    out = _ops().corr2d_lookup(_ops().corr2d_build(f1, f2, 4), coords, 4)
    g1, g2 = torch.autograd.grad(out, [f1, f2], gout)
    ref = R.corr2d_lookup(R.corr2d_build(f1, f2, 4), coords, 4)
    r1, r2 = torch.autograd.grad(ref, [f1, f2], gout)
    assert _rel(out, ref) <= 2e-5
    assert _rel(g1, r1) <= 2e-4 and _rel(g2, r2) <= 2e-4


def test_recompute_backward_operators(dev):
    """knn_interpolate, corr3d build + lookup, PointConv grouping, PointConvDW WeightNet, CLFM interpolation:
    kernel forward, formula-recompute backward."""
    from camliflow_b200.mlp import Conv2dNormRelu, MLP2d
    ops = _ops()
    g = torch.Generator().manual_seed(43)
    B, n, C = 1, 512, 32
    xyz1 = ((torch.rand(B, 3, n, generator=g) - 0.5) * 10).to(dev)
    xyz2 = (xyz1.cpu() + torch.randn(B, 3, n, generator=g) * 0.1).to(dev)
    xyzs2 = [xyz2, xyz2[:, :, :256].contiguous(), xyz2[:, :, :128].contiguous()]
    f1 = torch.randn(B, C, n, generator=g).to(dev).requires_grad_(True)
    f2 = torch.randn(B, C, n, generator=g).to(dev).requires_grad_(True)
    cost = MLP2d(4, [32, 32], act="relu").to(dev)
    P = [cost.convs[0].conv_fn.weight.flatten(1), cost.convs[0].conv_fn.bias, cost.convs[1].conv_fn.weight.flatten(1),
         cost.convs[1].conv_fn.bias]

    pyr = ops.corr3d_build(f1, f2, xyzs2)
    out = ops.corr3d_lookup_rows(xyz1, xyzs2, pyr, *P)
    gout = torch.randn(out.shape, generator=g).to(dev)
    got = torch.autograd.grad(out, [f1, f2] + list(cost.parameters()), gout)
    rp = [torch.bmm(f1.transpose(1, 2), f2) / C]
    for i in (1, 2):
        rp.append(R.corr3d_pool(rp[-1], _knn(xyzs2[i - 1], xyzs2[i], 3)))
    ref = R.corr3d_lookup(xyz1, xyzs2, rp, [_knn(x, xyz1, 16) for x in xyzs2], *P).transpose(1, 2)
    want = torch.autograd.grad(ref, [f1, f2] + list(cost.parameters()), gout)
    assert _rel(out, ref) <= 2e-5
    for a, b in zip(got, want):
        assert _rel(a, b) <= 5e-4

    flow = torch.randn(B, 3, n, generator=g).to(dev).requires_grad_(True)
    q = ((torch.rand(B, 3, 900, generator=g) - 0.5) * 10).to(dev)
    up = ops.knn_interpolate(xyz1, flow, q, 3)
    gq = torch.randn(up.shape, generator=g).to(dev)
    assert _rel(torch.autograd.grad(up, flow, gq)[0],
                torch.autograd.grad(R.knn_interpolate(xyz1, flow, q, _knn(xyz1, q, 3)), flow, gq)[0]) <= 1e-5

    wn = MLP2d(3, [8, 16], act="leaky_relu").to(dev)
    feat = torch.randn(B, C, n, generator=g).to(dev).requires_grad_(True)
    S = 256
    sx = xyz1[:, :, :S].contiguous()
    idx = _knn(xyz1, sx, 16)
    rows = ops.rows_of(torch.cat([xyz1, feat], 1))
    grp = ops.pointconv_group(rows, sx, idx, 16, wn, 0.1)
    gg = torch.randn(grp.shape, generator=g).to(dev)
    got = torch.autograd.grad(grp, [feat] + list(wn.parameters()), gg)
    fp = [wn.convs[0].conv_fn.weight.flatten(1), wn.convs[0].conv_fn.bias, wn.convs[1].conv_fn.weight.flatten(1),
          wn.convs[1].conv_fn.bias]
    want = torch.autograd.grad(R.pointconv_group(xyz1, feat, sx, idx, *fp, 0.1), [feat] + list(wn.parameters()), gg)
    for a, b in zip(got, want):
        assert _rel(a, b) <= 5e-4

    wn3 = MLP2d(3, [8, 32, 40], act="relu").to(dev)
    w = ops.pointconv_dw_weights(xyz1, xyz1, _knn(xyz1, xyz1, 16), 16, wn3)
    gw = torch.randn(w.shape, generator=g).to(dev)
    got = torch.autograd.grad(w, list(wn3.parameters()), gw)
    fp = []
    for c in wn3.convs:
        fp += [c.conv_fn.weight.flatten(1), c.conv_fn.bias]
    want = torch.autograd.grad(R.pointconv_dw_weights(xyz1, xyz1, _knn(xyz1, xyz1, 16), fp), list(wn3.parameters()), gw)
    for a, b in zip(got, want):
        assert _rel(a, b) <= 5e-4

    sn = torch.nn.Sequential(Conv2dNormRelu(3, 16), Conv2dNormRelu(16, C, act="sigmoid")).to(dev)
    H, W = 12, 20
    uv = torch.stack([torch.rand(B, n, generator=g) * (W - 1), torch.rand(B, n, generator=g) * (H - 1)], 1).to(dev)
    nn_idx = ops.nearest_point_2d(uv, H, W)
    f3 = torch.randn(B, n, C, generator=g).to(dev)
    ci = ops.clfm_interp(uv, nn_idx, f3, sn, H, W)
    gc = torch.randn(ci.shape, generator=g).to(dev)
    got = torch.autograd.grad(ci, list(sn.parameters()), gc)
    fp = [sn[0].conv_fn.weight.flatten(1), sn[0].conv_fn.bias, sn[1].conv_fn.weight.flatten(1), sn[1].conv_fn.bias]
    want = torch.autograd.grad(R.clfm_interp(uv, nn_idx, f3.transpose(1, 2), *fp, H, W), list(sn.parameters()), gc)
    for a, b in zip(got, want):
        assert _rel(a, b) <= 5e-4


@contextlib.contextmanager
def formula_ops():
    """camliflow_b200.ops with every differentiable operator answered by its plain-PyTorch formula (the index
    searches stay on the CUDA kernels): the pure-autograd checker of the training step."""
    import camliflow_b200.ops as ops

    def folded(layers):
        out = []
        for c in layers:
            w, b = c.folded()
            out += [w, b]
        return out

    def knn_interpolate(input_xyz, input_feat, query_xyz, k=3):
        return R.knn_interpolate(input_xyz, input_feat, query_xyz, _knn(input_xyz, query_xyz, k))

    def corr3d_build(feat1, feat2, xyzs2, k=3):
        pyr = [torch.bmm(feat1.transpose(1, 2), feat2) / feat1.shape[1]]
        for i in range(1, len(xyzs2)):
            pyr.append(R.corr3d_pool(pyr[-1], _knn(xyzs2[i - 1], xyzs2[i], k)))
        return pyr

    patches = {
        "knn_interpolate": knn_interpolate,
        "backwarp_3d": lambda a, b, f, k=3: b + knn_interpolate(a + f, -f, b, k),
        "bilinear_sample_rows": lambda f, uv: ops.rows_of(R.bilinear_sample(f, uv)),
        "corr2d_build": R.corr2d_build,
        "corr2d_lookup": lambda pyr, c, r, channels_last=True: R.corr2d_lookup(pyr, c, r),
        "corr3d_build": corr3d_build,
        "corr3d_lookup_rows": lambda xyz1, xyzs2, pyr, W1, b1, W2, b2: R.corr3d_lookup(
            xyz1, xyzs2, pyr, [_knn(x, xyz1, 16) for x in xyzs2], W1, b1, W2, b2).transpose(1, 2).contiguous(),
        "pointconv_dw_weights": lambda xyz, s, idx, k, wn: R.pointconv_dw_weights(xyz, s, idx[:, :, :k], folded(wn.convs)),
        "pointconv_dw_gather_max": lambda f, w, idx, k: ops.rows_of(R.pointconv_dw_gather_max(ops.cf_of(f), w, idx[:, :, :k])),
        "pointconv_group": lambda rows, sx, idx, k, wn, slope: R.pointconv_group(
            ops.cf_of(rows)[:, :3], ops.cf_of(rows)[:, 3:], sx, idx[:, :, :k], *folded(wn.convs), slope),
        "clfm_interp": lambda uv, nn, f, sn, H, W: R.clfm_interp(uv, nn, ops.cf_of(f), *folded(sn), H, W),
    }
    saved = {k: getattr(ops, k) for k in patches}
    try:
        for k, fn in patches.items():
            setattr(ops, k, fn)
        yield
    finally:
        for k, fn in saved.items():
            setattr(ops, k, fn)


def test_camliraft_training_step_gradients(dev):
    """One training step (train mode, 3 refinement iterations, sequence losses, backward) through the fused
    operators against the same step with every operator replaced by its torch formula."""
    from camliflow_b200.camliraft import CamLiRAFT
    from camliflow_b200.config import camliraft_config
    from camliflow_b200.init import seed_module_
    from oracle import camliraft_oracle as co
    inputs = {k: v.to(dev) for k, v in co.synthetic_inputs(1, 128, 160, 8192, seed=17).items()}
    # (a smaller image makes the REFERENCE's own coarsest lookup divide by h - 1 = 0)
    g = torch.Generator().manual_seed(18)
    inputs["flow_2d"] = (torch.randn(1, 2, 128, 160, generator=g) * 3).to(dev)
    inputs["flow_3d"] = (torch.randn(1, 3, 8192, generator=g) * 0.1).to(dev)
    model = seed_module_(CamLiRAFT(camliraft_config(n_iters_train=3)), seed=0).to(dev).train()

    def step():
        model.zero_grad(set_to_none=True)
        model(inputs)
        model.loss.backward()
        return model.loss.item(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

    loss, grads = step()
    with formula_ops():
        ref_loss, ref_grads = step()
    assert abs(loss - ref_loss) <= 1e-4 * abs(ref_loss), (loss, ref_loss)
    assert set(grads) == set(ref_grads)
    # every trainable tensor of the model receives a gradient
    assert len(grads) == sum(1 for _ in model.parameters())
    worst = max((_rel(grads[n], ref_grads[n]), n) for n in grads if ref_grads[n].abs().max() > 1e-9)
    print("training step: loss %.6f (formula %.6f), worst relative gradient difference %.2e at %s" % ((loss, ref_loss) + worst))
    assert worst[0] <= 5e-3, worst


# ------------------------------------------------------------------ training parity against the REFERENCE's gradients
def _golden_train_step(model_ctor, npz, shape, dev):
    """One train-mode step of the product on the GPU (fused forward kernels, hand-written / recompute backwards)
    against loss and per-parameter gradients of the reference model (tests/golden/make_golden_r2.py)."""
    import os
    import numpy as np
    from oracle import camliraft_oracle as co
    from tests._util import GOLDEN
    from tests.test_train_golden import compare_with_golden, train_targets
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    G = np.load(os.path.join(GOLDEN, npz))
    H, W, N, B, seed, tseed = shape
    inputs = dict(co.synthetic_inputs(B, H, W, N, seed), **train_targets(B, H, W, N, tseed))
    inputs = {k: v.to(dev) for k, v in inputs.items()}
    model = model_ctor().to(dev).train()
    model(inputs)
    model.loss.backward()
    # loss: fp32 reassociation only.  Gradients: a neighbour that flips in one of the searches on a predicted
    # (flow-warped) cloud changes a handful of rows, so the norms carry a little more than rounding
    worst = compare_with_golden(model, G, "small", loss_rtol=2e-4, norm_rtol=2e-2, sample_tol=5e-2)
    print("%s train step on GPU vs reference golden: loss %.6f (ref %.6f), worst grad-norm diff %.2e at %s"
          % ((npz, float(model.loss.detach()), float(G["small_loss"])) + worst))
    return model


def test_camliraft_training_step_vs_reference_golden(dev):
    from camliflow_b200.camliraft import CamLiRAFT
    from camliflow_b200.config import camliraft_config
    from camliflow_b200.init import seed_module_
    model = _golden_train_step(lambda: seed_module_(CamLiRAFT(camliraft_config(n_iters_train=3)), seed=0),
                               "train_camliraft.npz", (160, 224, 8192, 2, 17, 18), dev)
    m = model.get_metrics()
    assert {"loss", "loss2d", "loss3d", "epe2d", "acc2d_1px", "outlier2d", "epe3d", "acc3d_5cm"} <= set(m)


def test_camlipwc_training_step_vs_reference_golden(dev):
    from camliflow_b200.camlipwc import CamLiPWC
    from camliflow_b200.config import camlipwc_config
    from camliflow_b200.init import seed_module_
    _golden_train_step(lambda: seed_module_(CamLiPWC(camlipwc_config()), seed=0),
                       "train_camlipwc.npz", (128, 192, 8192, 2, 21, 22), dev)


def test_captured_train_step_matches_eager(dev):
    """trainer.CapturedTrainStep: the whole step (forward, losses, backward, flat gradient buffer, clip, AdamW) replayed
    from ONE CUDA graph gives the losses of the same step run eagerly, over several consecutive steps (i.e. the
    captured optimizer really updates the weights the next replay reads)."""
    from camliflow_b200 import trainer
    from camliflow_b200.camliraft import CamLiRAFT
    from camliflow_b200.config import camliraft_config
    from camliflow_b200.init import seed_module_
    from oracle import camliraft_oracle as co
    from tests.test_train_golden import train_targets
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, H, W, N = 1, 128, 160, 8192
    inputs = dict(co.synthetic_inputs(B, H, W, N, 31), **train_targets(B, H, W, N, 32))
    inputs = {k: v.to(dev) for k, v in inputs.items()}
    out = {}
    for mode in (False, True):
        model = seed_module_(CamLiRAFT(camliraft_config(n_iters_train=2)), seed=0).to(dev).train()
        step = trainer.CapturedTrainStep(model, inputs, lr=1e-3, use_graph=mode, warmup=0 if not mode else 2)
        if mode:     # the warm-up steps already moved the weights: restart from the seeded ones, in place
            seed_module_(model, seed=0)
            for st in step.opt.state.values():
                for v in st.values():
                    v.zero_()
        out[mode] = [float(step(inputs)) for _ in range(3)]
    print("captured step losses", out[True], "eager", out[False])
    assert out[True][0] != out[True][2]                          # the weights move
    for a, b in zip(out[True], out[False]):
        assert abs(a - b) <= 2e-4 * abs(b), (out[True], out[False])


def test_fused_wgrad_accumulation_matches_autograd(dev):
    """grad.fused_wgrad_accumulation (the captured training step's backward): the weight-gradient kernel of every 1x1 layer
    adds straight into the parameter's slice of the flat gradient buffer (CAMLI_WGRAD_ACCUMULATE) instead of returning a
    tensor to autograd's AccumulateGrad -- same gradients (float atomics: 1e-5 of the buffer's norm), and it is really taken."""
    from camliflow_b200 import grad, ops, trainer
    from camliflow_b200.camliraft import CamLiRAFT
    from camliflow_b200.config import camliraft_config
    from camliflow_b200.init import seed_module_
    from oracle import camliraft_oracle as co
    from tests.test_train_golden import train_targets
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, H, W, N = 1, 128, 160, 8192
    inputs = dict(co.synthetic_inputs(B, H, W, N, 41), **train_targets(B, H, W, N, 42))
    inputs = {k: v.to(dev) for k, v in inputs.items()}
    model = seed_module_(CamLiRAFT(camliraft_config(n_iters_train=2)), seed=0).to(dev).train()
    model.track_metrics = False
    flat = trainer.flatten_gradients(model)
    taken = []
    real = ops.conv_wgrad

    def counting(*a, **kw):
        taken.append(kw.get("accumulate_into") is not None)
        return real(*a, **kw)

    def run(fused):
        flat.zero_()
        grad.clear_dense_cache()
        model(inputs)
        with grad.fused_wgrad_accumulation(fused):
            model.loss.backward()
        return flat.clone()

    ops.conv_wgrad = counting
    try:
        plain = run(False)
        n_plain = sum(taken)
        fused = run(True)
    finally:
        ops.conv_wgrad = real
    assert n_plain == 0 and sum(taken) > 20, (n_plain, sum(taken), len(taken))
    err = (fused - plain).norm().item() / plain.norm().item()
    print("fused weight-gradient accumulation: %d of %d launches in place, relative difference %.2e" % (sum(taken), len(taken) // 2, err))
    assert err <= 1e-5


@pytest.mark.parametrize("B,H,W,s,scale", [(1, 68, 120, 8, 0.25), (2, 17, 23, 8, 0.25), (2, 33, 60, 4, 1.0), (1, 5, 3, 4, 1.0)])
def test_convex_upsample_forward_backward(dev, B, H, W, s, scale):
    """camli_convex_upsample{,_backward} against the reference formula (oracle.convex_upsample = models/utils.py:191-204)
    and its autograd gradients; RAFT's factor 8 with the 0.25 mask scale and CamLiPWC's factor 4."""
    from camliflow_b200 import ops
    from oracle import camliraft_oracle as co
    g = torch.Generator().manual_seed(B * 1000 + H)
    flow = (torch.randn(B, 2, H, W, generator=g) * 3).to(dev)
    mask = (torch.randn(B, 9 * s * s, H, W, generator=g) * 4).to(dev).contiguous(memory_format=torch.channels_last)
    gout = torch.randn(B, 2, H * s, W * s, generator=g).to(dev)
    f1, m1 = flow.clone().requires_grad_(True), mask.clone().requires_grad_(True)
    want = co.convex_upsample(f1, (scale * m1).contiguous(), s)
    want.backward(gout)
    f2, m2 = flow.clone().requires_grad_(True), mask.clone().requires_grad_(True)
    got = ops.convex_upsample(f2, m2, s, scale)
    got.backward(gout)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())      # fp32: summation order only
    for a, b, name in ((f2.grad, f1.grad, "flow"), (m2.grad, m1.grad, "mask")):
        assert (a - b).abs().max().item() <= 2e-5 * max(1.0, b.abs().max().item()), name
    with torch.no_grad():                                                                  # the inference doorway
        assert torch.equal(ops.convex_upsample(flow, mask, s, scale), got.detach())


@pytest.mark.parametrize("B,N,S,C,k", [(2, 700, 300, 32, 16), (1, 2048, 1024, 96, 16), (2, 512, 203, 128, 9), (1, 300, 64, 250, 16)])
def test_pointconv_group_backward_kernel(dev, B, N, S, C, k):
    """camli_pointconv_group_backward against autograd through the reference formula (tests/torch_ref.pointconv_group =
    models/point_conv.py:56-66): gradients of the feature AND coordinate columns, the centroids and the four WeightNet
    parameters.  Scatter-adds are float atomics: 1e-4 relative."""
    from camliflow_b200.mlp import MLP2d
    ops = _ops()
    g = torch.Generator().manual_seed(N + C)
    xyz = ((torch.rand(B, 3, N, generator=g) - 0.5) * 6).to(dev).requires_grad_(True)
    feat = torch.randn(B, C, N, generator=g).to(dev).requires_grad_(True)
    centre = (xyz.detach()[:, :, :S] + 0.01).contiguous().requires_grad_(True)
    idx = _knn(xyz.detach(), centre.detach(), k + 3)                    # a wider table than k, as the encoders pass
    wn = MLP2d(3, [8, 16], act="leaky_relu").to(dev)
    fp = [wn.convs[0].conv_fn.weight.flatten(1), wn.convs[0].conv_fn.bias, wn.convs[1].conv_fn.weight.flatten(1),
          wn.convs[1].conv_fn.bias]
    out = ops.pointconv_group(ops.rows_of(torch.cat([xyz, feat], 1)), centre, idx, k, wn, 0.1)
    gg = torch.randn(out.shape, generator=g).to(dev)
    leaves = [xyz, feat, centre] + list(wn.parameters())
    got = torch.autograd.grad(out, leaves, gg)
    want = torch.autograd.grad(R.pointconv_group(xyz, feat, centre, idx[:, :, :k], *fp, 0.1), leaves, gg)
    for name, a, b in zip(["xyz", "feat", "centre", "w1", "b1", "w2", "b2"], got, want):
        assert a.shape == b.shape and _rel(a, b) <= 1e-4, (name, _rel(a, b))
