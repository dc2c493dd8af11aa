"""Diagnosis of the 32-iteration case (C4): (1) every pair of the batch alone (B=1) vs in the batch of 4;
(2) per-iteration divergence of the product (GPU) from the CPU oracle on one pair."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from camliflow_b200.camliraft import CamLiRAFT  # noqa: E402
from camliflow_b200.config import camliraft_config  # noqa: E402
from camliflow_b200.init import seed_module_  # noqa: E402
from oracle import camliraft_oracle as co  # noqa: E402


def epe(a, b):
    return float((a - b).pow(2).sum(-2 if a.dim() == 2 else 1).sqrt().mean()) if False else float(np.sqrt(((a - b) ** 2).sum(0)).mean())


def main():
    pair = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    G = np.load(os.path.join(ROOT, "tests", "golden", "model_camliraft_c4.npz"))
    gain = json.loads(str(G["meta"]))["head_gain"]
    inputs = co.synthetic_inputs(4, 540, 960, 8192, seed=4)
    model = seed_module_(CamLiRAFT(camliraft_config(n_iters_eval=32)), seed=0).cuda().eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("flow_head.conv2.weight") or n.endswith("flow_head.fc.weight"):
                p.mul_(gain)
        out4 = model({k: v.cuda() for k, v in inputs.items()})
        for b in range(4):
            one = model({k: v[b:b + 1].cuda() for k, v in inputs.items()})
            e_b1 = epe(one["flow_2d"][0, :, ::8, ::8].cpu().numpy(), G["flow2d"][b]), epe(one["flow_3d"][0, :, ::4].cpu().numpy(), G["flow3d"][b])
            e_b4 = epe(out4["flow_2d"][b, :, ::8, ::8].cpu().numpy(), G["flow2d"][b]), epe(out4["flow_3d"][b, :, ::4].cpu().numpy(), G["flow3d"][b])
            same = float((one["flow_2d"][0] - out4["flow_2d"][b]).abs().max())
            print("pair %d: alone EPE2D %.2e EPE3D %.2e | in batch %.2e %.2e | max |alone - batch| %.2e" % ((b,) + e_b1 + e_b4 + (same,)), flush=True)
        # per-iteration trace of one pair against the CPU oracle
        model.core.all_predictions = True
        one_in = {k: v[pair:pair + 1] for k, v in inputs.items()}
        p2, p3 = model.predictions({k: v.cuda() for k, v in one_in.items()})
    P = co.make_params(co.param_spec("camliraft"), seed=0)
    for k in P:
        if k.endswith("flow_head.conv2.weight") or k.endswith("flow_head.fc.weight"):
            P[k] = P[k] * gain
    torch.set_num_threads(os.cpu_count() or 8)
    ref = co.camliraft_forward(P, one_in["images"], one_in["pcs"], one_in["intrinsics"], n_iters=32, index_impl="kernel", all_iters=True)
    r2, r3 = ref["flow_2d_preds"], ref["flow_3d_preds"]
    print("oracle final vs golden pair %d: EPE2D %.2e EPE3D %.2e" % (pair, epe(r2[-1][0, :, ::8, ::8].numpy(), G["flow2d"][pair]),
                                                                      epe(r3[-1][0, :, ::4].numpy(), G["flow3d"][pair])))
    for it in range(32):
        a2, a3 = p2[it][0].cpu().numpy(), p3[it][0].cpu().numpy()
        b2, b3 = r2[it][0].numpy(), r3[it][0].numpy()
        d2 = np.sqrt(((a2 - b2) ** 2).sum(0))
        d3 = np.sqrt(((a3 - b3) ** 2).sum(0))
        print("it %2d: EPE2D %.2e (max %.2e, >1e-2: %d px) EPE3D %.2e (max %.2e, >1e-3: %d pts)  |flow2d| %.2f" %
              (it, d2.mean(), d2.max(), int((d2 > 1e-2).sum()), d3.mean(), d3.max(), int((d3 > 1e-3).sum()),
               float(np.sqrt((b2 ** 2).sum(0)).mean())), flush=True)


if __name__ == "__main__":
    main()
