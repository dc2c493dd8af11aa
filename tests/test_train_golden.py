"""Training parity against the REFERENCE: loss and per-parameter gradients of one training step of the reference
models (tests/golden/train_camliraft.npz / train_camlipwc.npz, written by tests/golden/make_golden_r2.py from
/root/reference in train mode -- batch-statistics BatchNorm, sequence / pyramid losses, backward).

* CPU (this file, not gpu-marked): the product's module graph in train mode with the kernels answered by their
  formulas (tests/_cpu_ops) -- wiring, detach placement, BatchNorm modes, losses;
* GPU (tests/test_gpu_grad.py::test_training_step_vs_reference_golden*): the fused kernels and their backward
  kernels against the same fixture."""
import os

import numpy as np
import pytest
import torch

from oracle import camliraft_oracle as co
from tests._util import GOLDEN


def train_targets(B, H, W, N, seed):
    g = torch.Generator().manual_seed(seed)
    return {"flow_2d": torch.randn(B, 2, H, W, generator=g) * 3.0, "flow_3d": torch.randn(B, 3, N, generator=g) * 0.1}


def grad_sample(g, n=16):
    flat = g.reshape(-1)
    step = max(1, flat.numel() // n)
    s = flat[::step][:n]
    return torch.cat([s, s.new_zeros(n - s.numel())])


def compare_with_golden(model, G, case, loss_rtol, norm_rtol, sample_tol):
    """Asserts loss and every parameter gradient (norm, sum-free strided sample) against the fixture.  Returns the
    worst relative norm difference (for the log)."""
    loss = float(model.loss.detach())
    ref_loss = float(G[case + "_loss"])
    assert abs(loss - ref_loss) <= loss_rtol * abs(ref_loss), (loss, ref_loss)
    names = [str(n) for n in G[case + "_names"]]
    params = dict(model.named_parameters())
    assert set(names) == {n for n, p in params.items() if p.grad is not None} == set(params)
    gmax = float(G[case + "_norm"].max())
    worst = (0.0, None)
    for i, n in enumerate(names):
        g = params[n].grad.detach().float().cpu()
        ref_norm = float(G[case + "_norm"][i])
        # gradients far below the largest one are compared on the absolute scale of the step
        denom = max(ref_norm, 1e-4 * gmax)
        d = abs(float(g.double().norm()) - ref_norm) / denom
        worst = max(worst, (d, n))
        s, rs = grad_sample(g).numpy(), G[case + "_sample"][i]
        scale = max(float(np.abs(rs).max()), ref_norm / max(1.0, g.numel() ** 0.5), 1e-12)
        assert float(np.abs(s - rs).max()) <= sample_tol * scale + 1e-4 * gmax / max(1.0, g.numel() ** 0.5), (n, s, rs)
    assert worst[0] <= norm_rtol, worst
    return worst


def test_camliraft_training_step_host_graph_vs_reference_golden():
    from camliflow_b200.camliraft import CamLiRAFT
    from camliflow_b200.config import camliraft_config
    from camliflow_b200.init import seed_module_
    from tests._cpu_ops import cpu_kernels
    G = np.load(os.path.join(GOLDEN, "train_camliraft.npz"))
    H, W, N, B, iters, seed, tseed = 160, 224, 8192, 2, 3, 17, 18
    inputs = dict(co.synthetic_inputs(B, H, W, N, seed), **train_targets(B, H, W, N, tseed))
    model = seed_module_(CamLiRAFT(camliraft_config(n_iters_train=iters)), seed=0).train()
    with cpu_kernels():
        model(inputs)
        model.loss.backward()
    worst = compare_with_golden(model, G, "small", loss_rtol=1e-5, norm_rtol=2e-3, sample_tol=5e-3)
    print("CamLiRAFT train step (host graph, CPU formulas) vs reference: loss %.6f, worst grad-norm diff %.2e at %s"
          % ((float(model.loss),) + worst))
    ref_metrics = eval(str(G["small_metrics"]).replace("NaN", "float('nan')"))
    got = model.get_metrics()
    for k, v in ref_metrics.items():
        assert abs(got[k] - v) <= 1e-4 * max(1.0, abs(v)), (k, got[k], v)


def test_camlipwc_training_step_host_graph_vs_reference_golden():
    from camliflow_b200.camlipwc import CamLiPWC
    from camliflow_b200.config import camlipwc_config
    from camliflow_b200.init import seed_module_
    from tests._cpu_ops import cpu_kernels
    G = np.load(os.path.join(GOLDEN, "train_camlipwc.npz"))
    H, W, N, B, seed, tseed = 128, 192, 8192, 2, 21, 22
    inputs = dict(co.synthetic_inputs(B, H, W, N, seed), **train_targets(B, H, W, N, tseed))
    model = seed_module_(CamLiPWC(camlipwc_config()), seed=0).train()
    with cpu_kernels():
        model(inputs)
        model.loss.backward()
    worst = compare_with_golden(model, G, "small", loss_rtol=1e-5, norm_rtol=2e-3, sample_tol=5e-3)
    print("CamLiPWC train step (host graph, CPU formulas) vs reference: loss %.6f, worst grad-norm diff %.2e at %s"
          % ((float(model.loss),) + worst))
