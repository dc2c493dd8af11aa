"""CPU tests of the product's host side (module wiring, layouts, caches, state_dict names) with the
kernels answered by their PyTorch formulas (tests/_cpu_ops.py), against the REFERENCE model's golden
outputs.  The kernels themselves are tested on the GPU (tests/test_gpu_ops.py, test_gpu_model.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import camliraft_oracle as co
from tests._cpu_ops import cpu_kernels
from tests._util import GOLDEN


def epe(a, b):
    return float(np.sqrt(((a - b) ** 2).sum(0)).mean())


def _model(n_iters):
    from camliflow_b200.camliraft import CamLiRAFT
    from camliflow_b200.config import camliraft_config
    from camliflow_b200.init import seed_module_
    return seed_module_(CamLiRAFT(camliraft_config(n_iters_eval=n_iters)), seed=0).eval()


def test_state_dict_is_the_reference_state_dict():
    sd = _model(1).state_dict()
    spec = co.param_spec("camliraft")
    assert set(sd) == set(spec)
    assert all(tuple(sd[k].shape) == spec[k] for k in spec)
    P = co.make_params(spec, seed=0)
    assert all(torch.equal(sd[k], P[k]) for k in spec)


def test_product_graph_matches_reference_golden_small():
    G = np.load(os.path.join(GOLDEN, "model_camliraft.npz"))
    inputs = co.synthetic_inputs(1, 160, 224, 8192, seed=11)
    with cpu_kernels(), torch.no_grad():
        out = _model(3)(inputs)
    e2 = epe(out["flow_2d"][0, :, ::4, ::4].numpy(), G["small_kernel_flow2d"])
    e3 = epe(out["flow_3d"][0, :, ::4].numpy(), G["small_kernel_flow3d"])
    assert e2 <= 1e-3 and e3 <= 1e-4, (e2, e3)


def test_pipelined_point_branch_is_the_same_computation(monkeypatch):
    """The software-pipelined schedule of the recurrent core (the point branch's back-warp + correlation lookup of
    iteration i+1 issued behind the point update of iteration i) only moves launches: bit-identical flows."""
    inputs = co.synthetic_inputs(1, 160, 224, 8192, seed=11)
    outs = {}
    for mode in ("0", "force"):
        monkeypatch.setenv("CAMLI_PIPELINE_3D", mode)
        with cpu_kernels(), torch.no_grad():
            outs[mode] = _model(3)(inputs)
    assert bool(torch.isfinite(outs["0"]["flow_3d"]).all())
    assert torch.equal(outs["0"]["flow_2d"], outs["force"]["flow_2d"])
    assert torch.equal(outs["0"]["flow_3d"], outs["force"]["flow_3d"])


def test_module_surface_matches_rows_fast_path():
    """PointConvDW / CLFM public (channel-first) calls equal their channel-last fast paths."""
    from camliflow_b200 import ops
    from camliflow_b200.clfm import CLFM
    from camliflow_b200.point_conv import PointConvDW
    g = torch.Generator().manual_seed(0)
    xyz = torch.rand(2, 3, 300, generator=g) * 4
    feat = torch.randn(2, 24, 300, generator=g)
    with cpu_kernels(), torch.no_grad():
        conv = PointConvDW(24, 40, k=8).eval()
        a = conv(xyz, feat)
        cache = {}
        b = conv.forward_rows(xyz, ops.rows_of(feat), cache=cache)
        c = conv.forward_rows(xyz, ops.rows_of(feat), cache=cache)      # served from the cache
        assert a.shape == (2, 40, 300) and torch.equal(a, b.transpose(1, 2)) and torch.equal(b, c)
        assert len(cache) == 1
        clfm = CLFM(16, 24, norm="batch_norm").eval()
        uv = torch.stack([torch.rand(2, 300, generator=g) * 11, torch.rand(2, 300, generator=g) * 8], 1)
        f2d = torch.randn(2, 16, 9, 12, generator=g)
        o2, o3 = clfm(uv, f2d, feat)
        assert o2.shape == f2d.shape and o3.shape == feat.shape


def test_ops_refuse_cpu_tensors():
    from camliflow_b200 import ops
    with pytest.raises(RuntimeError):
        ops.knn_interpolate(torch.rand(1, 3, 10), torch.rand(1, 3, 10), torch.rand(1, 3, 5))
    with pytest.raises(RuntimeError):
        ops.corr2d_lookup([torch.rand(1, 4, 2, 2)], torch.rand(1, 2, 2, 2), 4)


def test_camlipwc_state_dict_and_graph_match_reference_golden():
    """CamLiPWC (BASELINE config[2] model): reference state_dict names/shapes, and the product graph
    (kernels answered by formulas) against the reference model's golden output."""
    from camliflow_b200.camlipwc import CamLiPWC
    from camliflow_b200.config import camlipwc_config
    from camliflow_b200.init import seed_module_
    net = seed_module_(CamLiPWC(camlipwc_config()), seed=0).eval()
    spec = co.param_spec("camlipwc")
    sd = net.state_dict()
    assert set(sd) == set(spec) and all(tuple(sd[k].shape) == spec[k] for k in spec)
    P = co.make_params(spec, seed=0)
    assert all(torch.equal(sd[k], P[k]) for k in spec)
    G = np.load(os.path.join(GOLDEN, "model_camlipwc.npz"))
    inputs = co.synthetic_inputs(1, 128, 192, 8192, seed=21)
    with cpu_kernels(), torch.no_grad():
        out = net(inputs)
    e2 = epe(out["flow_2d"][0, :, ::4, ::4].numpy(), G["small_kernel_flow2d"])
    e3 = epe(out["flow_3d"][0, :, ::4].numpy(), G["small_kernel_flow3d"])
    assert e2 <= 1e-3 and e3 <= 1e-4, (e2, e3)


def test_camliraft_l_c1_on_cpu_matches_reference_golden():
    """BASELINE config 1 (the reference's own CPU-runnable case): CamLiRAFT-L, 8192-point pair -> 2048 working
    points, 4 GRU iterations, batch 1, on CPU (kernels answered by formulas): reference state_dict
    names/shapes and the reference model's golden output; plus a ragged batch of two 4500-point pairs."""
    from camliflow_b200.camliraft_l import CamLiRAFT_L
    from camliflow_b200.config import camliraft_l_config
    from camliflow_b200.init import seed_module_
    spec = co.param_spec("camliraft_l")
    G = np.load(os.path.join(GOLDEN, "model_camliraft_l.npz"))
    for case, B, N, iters, seed in (("c1", 1, 8192, 4, 21), ("c1_batch2", 2, 4500, 2, 22)):
        net = seed_module_(CamLiRAFT_L(camliraft_l_config(n_iters_eval=iters)), seed=0).eval()
        sd = net.state_dict()
        assert set(sd) == set(spec) and all(tuple(sd[k].shape) == spec[k] for k in spec)
        inputs = co.synthetic_inputs(B, 540, 960, N, seed)
        with cpu_kernels(), torch.no_grad():
            out = net({"pcs": inputs["pcs"], "intrinsics": inputs["intrinsics"]})
        want = G[case + "_kernel_flow3d"]
        for b in range(B):
            e3 = epe(out["flow_3d"][b, :, ::4].numpy(), want[b])
            assert e3 <= 1e-4, (case, b, e3)


def test_disp2pc_matches_reference_formula():
    """utils.disp2pc against the reference's numpy arithmetic (utils.py:319-339), with and without flow."""
    from camliflow_b200.utils import disp2pc
    rng = np.random.default_rng(0)
    disp = rng.uniform(1.0, 80.0, (37, 53)).astype(np.float32)
    flow = rng.normal(0, 3, (37, 53, 2)).astype(np.float32)
    f, cx, cy, baseline = 721.5, 609.5, 172.8, 0.54
    for fl in (None, flow):
        depth = baseline * f / (disp + 1e-5)
        xx = np.tile(np.arange(53, dtype=np.float32)[None, :], (37, 1))
        yy = np.tile(np.arange(37, dtype=np.float32)[:, None], (1, 53))
        if fl is not None:
            xx, yy = xx + fl[..., 0], yy + fl[..., 1]
        want = np.stack([(xx - cx) * depth / f, (yy - cy) * depth / f, depth], -1)
        got = disp2pc(torch.from_numpy(disp), baseline, f, cx, cy, None if fl is None else torch.from_numpy(fl)).numpy()
        assert np.allclose(got, want, rtol=1e-6, atol=1e-6)
