"""CPU unit tests of host-side machinery that needs no kernel: the layer doorway's torch route (tc.py), the
recompute-autograd wrapper (grad.py), the layout helper of the tensor-core kernel, and the round-robin order of
EnginePool."""
import pytest
import torch
import torch.nn as nn

from camliflow_b200 import grad, ops, tc


def test_tc_conv2d_and_linear_torch_route_equal_the_modules():
    """On CPU tensors tc.conv2d / tc.linear take their torch route: conv -> (eval BN) -> + residual -> activation."""
    g = torch.Generator().manual_seed(0)
    conv, bn = nn.Conv2d(8, 6, 3, padding=1), nn.BatchNorm2d(6).eval()
    bn.running_mean.copy_(torch.randn(6, generator=g)), bn.running_var.copy_(torch.rand(6, generator=g) + 0.5)
    x, r = torch.randn(2, 8, 5, 7, generator=g), torch.randn(2, 6, 5, 7, generator=g)
    with torch.no_grad():
        assert not tc.fused(x)
        want = torch.relu(bn(conv(x)) + r)
        assert torch.allclose(tc.conv2d(x, conv, "relu", bn=bn, residual=r), want, atol=1e-6)
        out = torch.zeros(2, 5, 7, 6)
        got = tc.conv2d(x, conv, "relu", bn=bn, residual=r, out=out)
        assert torch.allclose(got, want, atol=1e-6) and got.data_ptr() == out.data_ptr()
        lin = nn.Linear(8, 5)
        rows = torch.randn(3, 11, 8, generator=g)
        assert torch.allclose(tc.linear(rows, lin.weight, lin.bias, "leaky_relu"), nn.functional.leaky_relu(lin(rows), 0.1), atol=1e-6)
        bn1 = nn.BatchNorm1d(5).eval()
        bn1.running_var.copy_(torch.rand(5, generator=g) + 0.5)
        want = torch.tanh(bn1(lin(rows).transpose(1, 2)).transpose(1, 2))
        assert torch.allclose(tc.linear(rows, lin.weight, lin.bias, "tanh", bn=bn1), want, atol=1e-5)


def test_pixel_layout_accepts_channel_slices_and_rejects_misaligned_views():
    t = torch.zeros(2, 6, 8, 96)
    assert ops._pixel_layout(t) == (96, True)
    assert ops._pixel_layout(t[..., 32:]) == (96, True)                 # channel slice at a 128-byte offset
    assert ops._pixel_layout(t[..., 2:34])[1] is False                   # 8-byte offset: not 16-byte aligned
    assert ops._pixel_layout(torch.zeros(2, 6, 8, 30))[1] is False        # pixel stride not a multiple of 4 floats
    assert ops._pixel_layout(t[:, ::2])[1] is False                       # rows skipped: not a uniform pixel stride
    nchw = torch.zeros(2, 96, 6, 8).contiguous(memory_format=torch.channels_last)
    assert ops._pixel_layout(nchw.permute(0, 2, 3, 1)) == (96, True)


def test_recompute_wrapper_differentiates_through_the_formula():
    """grad.recompute: forward = the 'kernel' (here a no-grad stand-in), backward = autograd through the formula on
    detached copies; non-tensor arguments and tensors that need no gradient pass through."""
    calls = {"kernel": 0, "formula": 0}

    def formula(a, b, scale, idx):
        calls["formula"] += 1
        return (a * b).sum(-1) * scale + idx.float()

    def kernel(a, b, scale, idx):
        calls["kernel"] += 1
        assert not torch.is_grad_enabled()
        return formula(a, b, scale, idx)

    a = torch.randn(4, 3, requires_grad=True)
    b = torch.randn(4, 3)
    idx = torch.arange(4)
    out = grad.recompute(kernel, formula, a, b, 2.0, idx)
    assert out.requires_grad and calls == {"kernel": 1, "formula": 1}
    out.sum().backward()
    assert torch.allclose(a.grad, 2.0 * b) and b.grad is None
    with torch.no_grad():
        plain = grad.recompute(kernel, formula, a, b, 2.0, idx)        # no autograd node when nothing needs a gradient
    assert not plain.requires_grad
    # list outputs
    outs = grad.recompute(lambda x: [x * 2, x + 1], lambda x: [x * 2, x + 1], a)
    assert isinstance(outs, list) and len(outs) == 2
    (outs[0].sum() + outs[1].sum()).backward()


class _FakeEngine:
    """Stands in for FlowEngine in the ordering test: records what it was asked to do."""

    def __init__(self, tag, log):
        self.tag, self.log, self.host_out = tag, log, {"tag": tag, "batch": None}

    def load(self, inputs):
        self._cur = inputs

    def step(self):
        self.log.append((self.tag, self._cur))

    def fetch_async(self):
        cur, host_out = self._cur, self.host_out

        class _Ev:
            def synchronize(self_inner):
                host_out["batch"] = cur
        return _Ev()


def test_engine_pool_round_robin_order():
    from camliflow_b200.engine import EnginePool
    pool = EnginePool.__new__(EnginePool)
    log = []
    pool.engines = [_FakeEngine(i, log) for i in range(3)]
    got = [(out["tag"], out["batch"]) for out in pool.pipelined(range(8))]
    assert [b for _, b in got] == list(range(8))                         # results come back in submission order
    assert [t for t, _ in got] == [i % 3 for i in range(8)]              # engines used round-robin
    assert log == [(i % 3, i) for i in range(8)]
    assert list(pool.pipelined([])) == []


def test_losses_match_reference_formulas():
    """camliflow_b200.losses against the reference's arithmetic (models/losses.py:64-119), with a validity mask."""
    from camliflow_b200.config import AttrDict
    from camliflow_b200.losses import calc_sequence_loss_2d, calc_sequence_loss_3d
    g = torch.Generator().manual_seed(1)
    preds = [torch.randn(2, 2, 5, 6, generator=g) for _ in range(3)]
    target = torch.cat([torch.randn(2, 2, 5, 6, generator=g), (torch.rand(2, 1, 5, 6, generator=g) > 0.3).float()], 1)
    for order, fn in (("l2-norm", lambda d: torch.linalg.norm(d, dim=1)), ("l1", lambda d: d.abs().sum(1)),
                      ("robust", lambda d: torch.pow(d.abs().sum(1) + 0.01, 0.4))):
        want = sum(0.8 ** (3 - i - 1) * fn(p - target[:, :2])[target[:, 2] > 0].mean() for i, p in enumerate(preds))
        assert torch.allclose(calc_sequence_loss_2d(preds, target, AttrDict(gamma=0.8, order=order)), want)
    p3 = [torch.randn(2, 3, 40, generator=g) for _ in range(2)]
    t3 = torch.randn(2, 3, 40, generator=g)
    want = sum(0.8 ** (2 - i - 1) * torch.linalg.norm(p - t3, dim=1).mean() for i, p in enumerate(p3))
    assert torch.allclose(calc_sequence_loss_3d(p3, t3, AttrDict(gamma=0.8, order="l2-norm")), want)
    with pytest.raises(ValueError):
        calc_sequence_loss_3d(p3, t3, AttrDict(gamma=0.8, order="nope"))


def test_latency_tile_policy_keeps_a_layer_within_one_wave():
    """ops._latency_tile: the narrowest of 32 / 64 / 128 output columns with m_tiles * n_tiles <= 148 CTAs; 0 (= the library's
    automatic width, which also knows the 96-column tile) when the widest tile is the answer or nothing fits."""
    from camliflow_b200 import ops
    assert ops._latency_tile(1, 1, 2048, 128) == 32          # 16 pixel tiles x 4
    assert ops._latency_tile(1, 68, 120, 128) == 64          # 68 x 2 = 136 (32 columns would need two waves)
    assert ops._latency_tile(1, 68, 120, 256) == 0           # 68 x 2 at 128 columns: the automatic width
    assert ops._latency_tile(1, 68, 120, 192) == 0           # -> two 96-column tiles inside the library
    assert ops._latency_tile(1, 68, 120, 324) == 0           # more than one wave at every width
    assert ops._latency_tile(4, 68, 120, 128) == 0           # four pairs per forward fill the GPU anyway
    assert ops._latency_tile(1, 1, 2048, 16) == 0            # one 32-column tile is the widest there is
    assert ops._latency_tile(1, 17, 30, 64) == 32


def test_fork_join_helper_runs_inline_without_streams():
    """camliraft_core._TwoStreams with streams disabled (CPU, training): run() and fork() execute in program order and
    hand the results through."""
    from camliflow_b200.camliraft_core import _Joined, _TwoStreams
    order = []
    par = _TwoStreams(False)
    a, b = par.run(lambda: order.append("main") or 1, lambda: order.append("side") or 2)
    h = par.fork(lambda: order.append("fork") or 3)
    assert (a, b, h.join(), h.join()) == (1, 2, 3, 3) and order == ["main", "side", "fork"]
    assert isinstance(h, _Joined) and h.stream is None


def test_fused_wgrad_accumulation_context_is_scoped_and_declines_without_buffers():
    """grad.fused_wgrad_accumulation only applies inside the context and only to leaf fp32 parameters that already own a
    contiguous, 16-byte aligned gradient buffer of the right size."""
    from camliflow_b200 import grad
    w = torch.nn.Parameter(torch.randn(8, 4, 1, 1))
    assert grad._grad_slot(w, 8, 4) is None                              # outside the context
    with grad.fused_wgrad_accumulation():
        assert grad._grad_slot(w, 8, 4) is None                          # no .grad yet
        w.grad = torch.zeros_like(w)
        slot = grad._grad_slot(w, 8, 4)
        assert slot is not None and slot.shape == (8, 4) and slot.data_ptr() == w.grad.data_ptr()
        lin = torch.nn.Parameter(torch.randn(8, 4))
        lin.grad = torch.zeros_like(lin)
        assert grad._grad_slot(lin[:, :, None, None], 8, 4).data_ptr() == lin.grad.data_ptr()    # a Linear weight seen as 4-D
        assert grad._grad_slot(w * 2.0, 8, 4) is None                    # not a leaf (e.g. a folded BatchNorm)
        with grad.fused_wgrad_accumulation(False):
            assert grad._grad_slot(w, 8, 4) is None
        assert grad._grad_slot(w, 8, 4) is not None
    assert grad._grad_slot(w, 8, 4) is None
