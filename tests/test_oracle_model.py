"""Pins oracle/camliraft_oracle.py (the CPU restatement of the reference's CamLiRAFT forward)
against outputs of the REFERENCE model itself (tests/golden/model_camliraft.npz, written by
tests/golden/make_golden_model.py from /root/reference with the same name-seeded weights)."""
import os

import numpy as np
import pytest
import torch

from oracle import camliraft_oracle as co
from tests._util import GOLDEN

GOLD = np.load(os.path.join(GOLDEN, "model_camliraft.npz"))
# the tolerance north_star states for flows: EPE2D <= 1e-3 px, EPE3D <= 1e-4 m
TOL_EPE2D, TOL_EPE3D = 1e-3, 1e-4


def epe(a, b):
    """End-point error, eval_things.py:62,88: mean over positions of the L2 norm over channels."""
    return float(np.sqrt(((a - b) ** 2).sum(0)).mean())


@pytest.fixture(scope="module")
def params():
    return co.make_params(co.param_spec("camliraft"), seed=0)


def test_param_spec_is_the_reference_state_dict(params):
    spec = co.param_spec("camliraft")
    assert len(spec) == 604 and sum(int(np.prod(s)) for s in spec.values()) == 8403185
    assert all(tuple(params[k].shape) == tuple(spec[k]) for k in spec)


@pytest.mark.parametrize("mode", ["fallback", "kernel"])
def test_oracle_matches_reference_small(params, mode):
    inputs = co.synthetic_inputs(1, 160, 224, 8192, seed=11)
    out = co.camliraft_forward(params, inputs["images"], inputs["pcs"], inputs["intrinsics"], n_iters=3,
                               index_impl=mode)
    f2 = out["flow_2d"][0, :, ::4, ::4].numpy()
    f3 = out["flow_3d"][0, :, ::4].numpy()
    assert epe(f2, GOLD["small_%s_flow2d" % mode]) <= TOL_EPE2D
    assert epe(f3, GOLD["small_%s_flow3d" % mode]) <= TOL_EPE3D


def test_oracle_matches_reference_c2(params):
    """BASELINE config[1]: 960x540 + 8192 points, 12 iterations, kernel index semantics."""
    inputs = co.synthetic_inputs(1, 540, 960, 8192, seed=0)
    out = co.camliraft_forward(params, inputs["images"], inputs["pcs"], inputs["intrinsics"], n_iters=12,
                               index_impl="kernel")
    e2 = epe(out["flow_2d"][0, :, ::8, ::8].numpy(), GOLD["c2_kernel_flow2d"])
    e3 = epe(out["flow_3d"][0, :, ::4].numpy(), GOLD["c2_kernel_flow3d"])
    assert e2 <= TOL_EPE2D and e3 <= TOL_EPE3D, (e2, e3)
