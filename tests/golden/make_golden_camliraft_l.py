"""Runs the REFERENCE CamLiRAFT-L (LiDAR-only) model -- BASELINE config 1: 8192-point pair (2048 working
points), 4 GRU iterations, batch 1, CPU -- imported from /root/reference (build container only) on seeded inputs
with name-seeded weights, in both index semantics (see make_golden_model.py), and writes
tests/golden/model_camliraft_l.npz.

    python tests/golden/make_golden_camliraft_l.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402
from make_golden_model import patch_indices  # noqa: E402
from oracle import camliraft_oracle as co  # noqa: E402

CASES = {"c1": (8192, 4, 21), "c1_batch2": (4500, 2, 22)}      # name: (N points, iterations, seed)


def main():
    torch.set_num_threads(8)
    models = rh.load_reference()
    P = co.make_params(co.param_spec("camliraft_l"), seed=0)
    out = {}
    for name, (N, iters, seed) in CASES.items():
        net = models.camliraft_l.CamLiRAFT_L(rh.camliraft_l_cfg(n_iters=iters)).eval()
        print(name, net.load_state_dict(P, strict=True))
        B = 2 if name.endswith("batch2") else 1
        inputs = co.synthetic_inputs(B, 540, 960, N, seed)
        inputs = {"pcs": inputs["pcs"], "intrinsics": inputs["intrinsics"]}
        for mode in ("fallback", "kernel"):
            patch_indices(models, mode == "kernel")
            with torch.no_grad():
                f3 = net(inputs)["flow_3d"]
            print(name, mode, tuple(f3.shape), float(f3.abs().mean()), float(f3.abs().max()))
            out["%s_%s_flow3d" % (name, mode)] = f3[:, :, ::4].numpy().astype(np.float32)
        patch_indices(models, False)
    np.savez_compressed(os.path.join(HERE, "model_camliraft_l.npz"), **out)


if __name__ == "__main__":
    main()
