"""The INTEGRATION.md section-1 injection, executed: the REFERENCE's own Python code (imported from
/root/reference, so build container only -- skipped on the GPU box) runs with `models.csrc` replaced by
`camliflow_b200.csrc` and `models.point_conv.PointConv{,DW}` / `models.clfm.CLFM` replaced by the product's
modules.  There is no GPU here, so the native entry points behind the product's Python surface are answered on
the CPU: the three `_..._cuda` doorways of camliflow_b200/csrc/wrapper.py by the C oracle
(oracle/kernels_oracle.c) and the fused operators by their formulas (tests/_cpu_ops.cpu_kernels).  What is under
test is everything between the reference's call sites and the C ABI: names, argument order, accepted layouts
(channel-first vs channel-last clouds, NCHW -> NHWC correlation inputs), index dtypes, the autograd Function of
the cost volume, module constructor signatures and state_dict keys.

Also here: the product's losses and metric bookkeeping against the reference's formulas."""
import contextlib
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import camliraft_oracle as co
from tests import _util
from tests._cpu_ops import cpu_kernels
from tests._util import GOLDEN

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="needs the reference checkout")


def epe(a, b):
    return float(np.sqrt(((a - b) ** 2).sum(0)).mean())


def _cpu_native():
    """CPU answers (C oracle) for the native doorways of camliflow_b200.csrc.wrapper."""
    def fps(points_xyz, n_samples):
        return torch.from_numpy(_util.oracle_fps(points_xyz.detach().numpy(), n_samples))

    def knn_strided(input_view, query_view, k):
        return torch.from_numpy(_util.oracle_knn(input_view.detach().contiguous().numpy(),
                                                 query_view.detach().contiguous().numpy(), k))

    def corr_fwd(in1, in2, md):
        return torch.from_numpy(_util.oracle_corr_fwd(in1.detach().numpy(), in2.detach().numpy(), md))

    def corr_bwd(gout, in1, in2, md):
        g1, g2 = _util.oracle_corr_bwd(gout.detach().contiguous().numpy(), in1.detach().numpy(), in2.detach().numpy(), md)
        return torch.from_numpy(g1), torch.from_numpy(g2)

    return {"_furthest_point_sampling_cuda": fps, "_k_nearest_neighbor_strided": knn_strided,
            "_k_nearest_neighbor_cuda": lambda i, q, k: knn_strided(i, q, k),
            "_correlation_forward_cuda": corr_fwd, "_correlation_backward_cuda": corr_bwd}


@contextlib.contextmanager
def injected_reference(swap_modules):
    """`import models` (the reference) with models.csrc = camliflow_b200.csrc; optionally the fused modules too."""
    sys.path.insert(0, os.path.join(GOLDEN))
    import ref_harness as rh
    import camliflow_b200.csrc as fast_csrc
    from camliflow_b200.csrc import wrapper
    saved_native = {k: getattr(wrapper, k) for k in _cpu_native()}
    saved_modules = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    for k in saved_modules:
        del sys.modules[k]
    try:
        for k, fn in _cpu_native().items():
            setattr(wrapper, k, fn)
        rh._install_stubs()
        if REF not in sys.path:
            sys.path.insert(0, REF)
        sys.modules["models.csrc"] = fast_csrc                       # INTEGRATION.md section 1, line 1
        models = importlib.import_module("models")
        assert sys.modules["models.utils"].k_nearest_neighbor is fast_csrc.k_nearest_neighbor
        assert sys.modules["models.pwc_core"].correlation2d is fast_csrc.correlation2d
        assert sys.modules["models.camlipwc_core"].k_nearest_neighbor is fast_csrc.k_nearest_neighbor
        if swap_modules:                                             # INTEGRATION.md section 1, second block
            import camliflow_b200.clfm as cf
            import camliflow_b200.point_conv as pc
            for name in ("models.point_conv", "models.camliraft_l_core", "models.camlipwc_l_core"):
                mod = sys.modules[name]
                mod.PointConv, mod.PointConvDW = pc.PointConv, pc.PointConvDW
            sys.modules["models.clfm"].CLFM = cf.CLFM
            sys.modules["models.camliraft_core"].CLFM = cf.CLFM
            sys.modules["models.camlipwc_core"].CLFM = cf.CLFM
        yield models, rh
    finally:
        for k, fn in saved_native.items():
            setattr(wrapper, k, fn)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved_modules)


def test_reference_camliraft_runs_on_injected_csrc():
    """The reference's CamLiRAFT (its own cores, point_conv, clfm) over the product's csrc: the golden output of the
    unmodified reference (kernel index semantics) is reproduced."""
    G = np.load(os.path.join(GOLDEN, "model_camliraft.npz"))
    inputs = co.synthetic_inputs(1, 160, 224, 8192, seed=11)
    with injected_reference(swap_modules=False) as (models, rh):
        net = models.camliraft.CamLiRAFT(rh.camliraft_cfg(n_iters=3)).eval()
        net.load_state_dict(co.make_params(co.param_spec("camliraft"), seed=0), strict=True)
        with torch.no_grad():
            out = net(inputs)
    e2 = epe(out["flow_2d"][0, :, ::4, ::4].numpy(), G["small_kernel_flow2d"])
    e3 = epe(out["flow_3d"][0, :, ::4].numpy(), G["small_kernel_flow3d"])
    assert e2 <= 1e-5 and e3 <= 1e-6, (e2, e3)


def test_reference_cores_accept_swapped_pointconv_and_clfm():
    """The reference's camliraft_core.py / camliraft_l_core.py schedulers over the product's PointConv,
    PointConvDW and CLFM modules (constructor signatures, forward signatures, state_dict keys)."""
    G = np.load(os.path.join(GOLDEN, "model_camliraft.npz"))
    inputs = co.synthetic_inputs(1, 160, 224, 8192, seed=11)
    with injected_reference(swap_modules=True) as (models, rh), cpu_kernels():
        import camliflow_b200.point_conv as pc
        net = models.camliraft.CamLiRAFT(rh.camliraft_cfg(n_iters=3)).eval()
        assert isinstance(net.core.branch_3d.fnet.convs[0], pc.PointConv)
        assert isinstance(net.core.branch_3d.gru.conv_z, pc.PointConvDW)
        net.load_state_dict(co.make_params(co.param_spec("camliraft"), seed=0), strict=True)   # same keys and shapes
        with torch.no_grad():
            out = net(inputs)
    e2 = epe(out["flow_2d"][0, :, ::4, ::4].numpy(), G["small_kernel_flow2d"])
    e3 = epe(out["flow_3d"][0, :, ::4].numpy(), G["small_kernel_flow3d"])
    assert e2 <= 1e-3 and e3 <= 1e-4, (e2, e3)


def test_reference_camlipwc_runs_on_injected_csrc_with_backward():
    """The reference's CamLiPWC over the product's csrc -- correlation2d included (models/pwc_core.py:205,
    models/camlipwc_core.py:182), forward against the golden and one backward through CorrelationFunction."""
    G = np.load(os.path.join(GOLDEN, "model_camlipwc.npz"))
    inputs = co.synthetic_inputs(1, 128, 192, 8192, seed=21)
    with injected_reference(swap_modules=False) as (models, rh):
        net = models.camlipwc.CamLiPWC(rh.camlipwc_cfg()).eval()
        net.load_state_dict(co.make_params(co.param_spec("camlipwc"), seed=0), strict=True)
        with torch.no_grad():
            out = net(inputs)
        e2 = epe(out["flow_2d"][0, :, ::4, ::4].numpy(), G["small_kernel_flow2d"])
        e3 = epe(out["flow_3d"][0, :, ::4].numpy(), G["small_kernel_flow3d"])
        assert e2 <= 1e-4 and e3 <= 1e-5, (e2, e3)
        # gradient through the drop-in cost volume
        csrc = sys.modules["models.csrc"]
        g = torch.Generator().manual_seed(0)
        a = torch.randn(1, 8, 6, 7, generator=g, requires_grad=True)
        b = torch.randn(1, 8, 6, 7, generator=g, requires_grad=True)
        csrc.correlation2d(a, b, 2).square().sum().backward()
        from tests import torch_ref as R
        a2, b2 = a.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
        R.correlation2d(a2, b2, 2).square().sum().backward()
        assert torch.allclose(a.grad, a2.grad, atol=1e-5) and torch.allclose(b.grad, b2.grad, atol=1e-5)


# ------------------------------------------------------------------------------------ losses and metrics
def _ref_module(name):
    sys.path.insert(0, os.path.join(GOLDEN))
    import ref_harness as rh
    rh.load_reference()
    return importlib.import_module(name)


@pytest.mark.parametrize("order", ["l2-norm", "robust"])
@pytest.mark.parametrize("sparse", [False, True])
def test_pyramid_losses_equal_reference(order, sparse):
    from camliflow_b200 import losses
    from camliflow_b200.config import AttrDict
    ref = _ref_module("models.losses")
    g = torch.Generator().manual_seed(3)
    cfg = AttrDict(level_weights=[8, 4, 2, 1, 0.5], order=order)
    B, H, W, N = 2, 64, 96, 1024
    flows2d = [torch.randn(B, 2, H >> (l + 2), W >> (l + 2), generator=g) for l in range(5)]
    t2 = torch.randn(B, 2, H, W, generator=g) * 3
    t3 = torch.randn(B, 3, N, generator=g)
    if sparse:
        t2 = torch.cat([t2, (torch.rand(B, 1, H, W, generator=g) > 0.4).float()], 1)
        t3 = torch.cat([t3, (torch.rand(B, 1, N, generator=g) > 0.4).float()], 1)
    sizes = [N, 512, 256, 128, 64]
    idx = [torch.stack([torch.randperm(N, generator=g)[:n] for _ in range(B)]) for n in sizes]
    flows3d = [torch.randn(B, 3, n, generator=g) for n in sizes]
    a = losses.calc_pyramid_loss_2d(flows2d, t2, cfg)
    b = ref.calc_pyramid_loss_2d([f.clone() for f in flows2d], t2, cfg)
    assert torch.allclose(a, b, rtol=1e-6), (a, b)
    import camliflow_b200.utils as ut
    saved = ut.batch_indexing
    ut.batch_indexing = lambda d, i, layout="channel_first": torch.gather(d, 2, i[:, None, :].expand(-1, d.shape[1], -1))
    try:
        a = losses.calc_pyramid_loss_3d(flows3d, t3, cfg, idx)
    finally:
        ut.batch_indexing = saved
    b = ref.calc_pyramid_loss_3d(flows3d, t3, cfg, idx)
    assert torch.allclose(a, b, rtol=1e-6), (a, b)


@pytest.mark.parametrize("order", ["l2-norm", "l1", "robust"])
def test_sequence_losses_equal_reference(order):
    from camliflow_b200 import losses
    from camliflow_b200.config import AttrDict
    ref = _ref_module("models.losses")
    g = torch.Generator().manual_seed(4)
    cfg = AttrDict(gamma=0.8, order=order)
    p2 = [torch.randn(2, 2, 24, 32, generator=g) for _ in range(4)]
    p3 = [torch.randn(2, 3, 500, generator=g) for _ in range(4)]
    t2 = torch.cat([torch.randn(2, 2, 24, 32, generator=g), (torch.rand(2, 1, 24, 32, generator=g) > 0.5).float()], 1)
    t3 = torch.cat([torch.randn(2, 3, 500, generator=g), (torch.rand(2, 1, 500, generator=g) > 0.5).float()], 1)
    assert torch.allclose(losses.calc_sequence_loss_2d(p2, t2, cfg), ref.calc_sequence_loss_2d(p2, t2, cfg), rtol=1e-6)
    assert torch.allclose(losses.calc_sequence_loss_3d(p3, t3, cfg), ref.calc_sequence_loss_3d(p3, t3, cfg), rtol=1e-6)
    assert torch.allclose(losses.calc_sequence_loss_2d(p2, t2[:, :2], cfg), ref.calc_sequence_loss_2d(p2, t2[:, :2], cfg), rtol=1e-6)


def test_metric_bookkeeping_equals_reference():
    """update_2d_metrics / update_3d_metrics / update_metrics / get_metrics over several calls, dense and sparse
    targets, occlusion mask -- against models/base.py (which syncs per metric; ours reduces once at read-out)."""
    from camliflow_b200.base import FlowModel
    ref = _ref_module("models.base")
    ours, theirs = FlowModel(), ref.FlowModel()
    g = torch.Generator().manual_seed(5)
    for step in range(3):
        p2, t2 = torch.randn(2, 2, 20, 30, generator=g) * 4, torch.randn(2, 2, 20, 30, generator=g) * 4
        p3, t3 = torch.randn(2, 3, 400, generator=g) * 0.1, torch.randn(2, 3, 400, generator=g) * 0.1
        if step == 1:
            t2 = torch.cat([t2, (torch.rand(2, 1, 20, 30, generator=g) > 0.5).float()], 1)
            t3 = torch.cat([t3, (torch.rand(2, 1, 400, generator=g) > 0.5).float()], 1)
        occ = (torch.rand(2, 400, generator=g) > 0.7).float()
        for m in (ours, theirs):
            m.update_metrics("loss", torch.tensor(1.5 + step))
            m.update_2d_metrics(p2, t2)
            m.update_3d_metrics(p3, t3)
            m.update_3d_metrics(p3, t3, occ)
    a, b = ours.get_metrics(), theirs.get_metrics()
    assert set(a) == set(b)
    for k in a:
        assert abs(a[k] - b[k]) <= 1e-6 * max(1.0, abs(b[k])), (k, a[k], b[k])
    ours.clear_metrics()
    assert ours.get_metrics() == {}
    with pytest.raises(ValueError):
        FlowModel().get_loss()
